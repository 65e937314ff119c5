"""Host-side compiler of the fused neck kernel (csrc/neck_fused.cu, C ABI ``yad_neck_fused``).

``MultiScaleFmapModule.forward`` (reference: modules/_common.py:241-265) at H = 1, deploy form, becomes a PROGRAM: a list of
ops (convolutions as lists of K blocks over shared-memory activation planes or over the backbone maps in global memory; the
max-pool cascade; bilinear x2 / x0.5; the even / odd split in front of the stride-2 convs) plus a shared-memory plan (plane
offsets from a first-fit allocator with explicit lifetimes) plus one weight blob in the order the kernel consumes it.  The
kernel interprets the program once per clip; this module only decides WHAT runs and WHERE it lives.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib

CONV, POOLS, PAIRAVG, UP2, DEINT, DUMP = 0, 1, 2, 3, 4, 5
SLOT = 16384
SMEM_MAX = 227 * 1024
ACT_LRELU = _lib.ACT_LRELU


def _ceil(a: int, b: int) -> int:
    return (a + b - 1) // b * b


TOKEN = 1 << 24      # plane handles are TOKEN + id until _finish() has placed them


class _Pool:
    """Shared-memory planner over 1024-byte units (planes start on a swizzle-atom boundary).  While the program is emitted a
    plane is only a handle with a lifetime [born, died) in op indices; place() then packs all planes offline: largest first, each
    at the lowest offset that is free during its whole lifetime (an online first-fit allocator needed 123 KB for a 90 KB peak)."""

    def __init__(self):
        self.planes: List[List[int]] = []      # [units, born, died]
        self.live: set = set()

    def alloc(self, nbytes: int, now: int) -> int:
        self.planes.append([_ceil(nbytes, 1024) // 1024, now, -1])
        self.live.add(len(self.planes) - 1)
        return TOKEN + len(self.planes) - 1

    def free(self, tok: int, now: int) -> None:
        i = tok - TOKEN
        self.live.remove(i)
        self.planes[i][2] = now

    def _place_order(self, order) -> Tuple[List[int], int]:
        off = [-1] * len(self.planes)
        for i in order:
            n, b, d = self.planes[i]
            busy = sorted((off[j], off[j] + self.planes[j][0]) for j in range(len(self.planes))
                          if off[j] >= 0 and self.planes[j][1] < d and b < self.planes[j][2])
            pos = 0
            for lo, hi in busy:
                if lo - pos >= n:
                    break
                pos = max(pos, hi)
            off[i] = pos
        high = max(off[i] + self.planes[i][0] for i in range(len(self.planes)))
        return off, high

    def place(self) -> Tuple[List[int], int]:
        """First-fit over several plane orders (largest first, by birth, by death, and seeded shuffles of the largest-first order):
        the best packing found is kept (136 -> ~122 KB on the two-clip plan of a 60 s clip, whose liveness peak is 120 KB)."""
        import random
        idx = list(range(len(self.planes)))
        orders = [sorted(idx, key=lambda i: (-self.planes[i][0], self.planes[i][1])),
                  sorted(idx, key=lambda i: (self.planes[i][1], -self.planes[i][0])),
                  sorted(idx, key=lambda i: (-(self.planes[i][2] - self.planes[i][1]), -self.planes[i][0])),
                  sorted(idx, key=lambda i: (-self.planes[i][0] * (self.planes[i][2] - self.planes[i][1]), self.planes[i][1]))]
        rng = random.Random(1234)
        for _ in range(400):
            o = list(orders[0])
            for _ in range(rng.randint(1, 6)):          # a few transpositions of the largest-first order
                a, b = rng.randrange(len(o)), rng.randrange(len(o))
                o[a], o[b] = o[b], o[a]
            orders.append(o)
        best = None
        for o in orders:
            off, high = self._place_order(o)
            if best is None or high < best[1]:
                best = (off, high)
        off, high = best
        return [o * 1024 for o in off], high * 1024


class FusedNeck:
    """Compiled neck for one input geometry.  ``eng`` is the InferenceEngine (packed convolutions ``eng.n`` / ``eng.rep``)."""

    def __init__(self, eng, Hs: Sequence[int], Ws: Sequence[int], Cs: Sequence[int], debug: bool = False, G: Optional[int] = None):
        self.eng, self.dev = eng, eng.dev
        self.Hs, self.Ws, self.Cs = list(Hs), list(Ws), list(Cs)
        W1, W2, W3, W4 = Ws
        if not (W1 == 2 * W2 and W2 == 2 * W3 and W3 == 2 * W4):
            raise ValueError("feature-map widths do not nest")
        for h in Hs:
            if h & (h - 1):
                raise NotImplementedError("fused neck: feature-map heights must be powers of two (exact 1/H in bf16)")
        for nm, blocks in eng.rep.items():
            if any("deploy" not in b for b in blocks) or len(blocks) != 2:
                raise NotImplementedError("fused neck: re-parameterised (deploy) RepBlocks of two blocks only")
        self.debug = debug
        if G is None and os.environ.get("YAD_NECK_G"):
            G = int(os.environ["YAD_NECK_G"])
        last = None
        for g in ([G] if G else [2, 1]):
            try:
                self._compile(g)
                return
            except NotImplementedError as e:
                last = e
        raise last

    def _compile(self, G: int):
        """G clips per CTA pass ("unit").  G = 1: clip pitch = W + 1 rows (the flat layout of the backbone).  G = 2: the clips of a
        unit sit at a power-of-two row pitch P (P4 >= W4 + 1, P3 = 2 P4, ...), so that at levels 3 / 4 BOTH clips share one 128-row
        M tile - the second clip costs no MMA and no epilogue pass there, and every weight block is fetched once per unit - and
        the row pairs of the x0.5 / even-odd folds stay inside a clip.  Rows w in [W, P) of a clip are zero (halo + padding)."""
        Ws = self.Ws
        self.G = G                                              # clips per CTA pass
        self.lv = {}                                            # level -> geometry
        p4 = Ws[3] + 1
        if G > 1:
            p4 = max(8, 1 << (p4 - 1).bit_length())         # >= one swizzle atom: a clip's TMA box starts on a 1024-byte boundary
        for i, W in enumerate(Ws):
            P = W + 1 if G == 1 else p4 << (3 - i)
            R = G * P
            self.lv[i + 1] = {"W": W, "Wp": P, "R": R, "n_mt": (R + 127) // 128, "bytes": _ceil(R + 2, 8) * 128}
            if self.lv[i + 1]["n_mt"] > 4:
                raise NotImplementedError("fused neck: clip too long (more than 4 M tiles per level)")
        self.pool = _Pool()
        self.ops: List[List[int]] = []
        self.kbs: List[Tuple[int, int, bool]] = []             # (plane / chunk, tap shift, global-sourced)
        self.wblocks: List[torch.Tensor] = []
        self.wrows = 0
        self.biases: List[torch.Tensor] = []
        self.nbias = 0
        self.dumps: Dict[str, Tuple[int, int, int]] = {}        # name -> (element offset, rows, level)
        self.dump_elems = 0
        self._build()
        self._finish()

    # ------------------------------------------------------------------ weights
    @staticmethod
    def _w4(cv) -> torch.Tensor:
        return cv.w.view(cv.cout_pad, cv.kh, cv.kw, cv.cin_pad).float()

    def _blk(self, cvs, kh: int, kw: int, chunk: int, scale: float = 1.0) -> torch.Tensor:
        """[N, 64] weight block of K chunk ``chunk`` at tap (kh, kw); several convs are stacked along N (merged convolutions)."""
        return torch.cat([self._w4(cv)[:, kh, kw, 64 * chunk:64 * chunk + 64] * scale for cv in cvs], 0)

    # ------------------------------------------------------------------ program emission
    def _plane(self, level: int) -> int:
        return self.pool.alloc(self.lv[level]["bytes"], len(self.ops))

    def _free(self, tok: int) -> None:
        self.pool.free(tok, len(self.ops))

    def _conv(self, cvs, level: int, kblocks, outs: List[int], head: int = -1, src_global: int = -1, pair: Sequence[bool] = (),
              deint: bool = False, no_epilogue: bool = False, accumulate: bool = False, row_base: int = 0, n_mt: Optional[int] = None):
        """kblocks: list of (src, shift, weight block [N, 64]).  pair[j]: output plane j is written pair-averaged (bilinear x0.5,
        F.interpolate in BiC's conv_c0 branch, modules/_common.py:181-182) at the NEXT level's geometry.  deint: outs = (even, odd)
        planes at the next level's geometry - the column split that the stride-(1, 2) conv behind it reads (N <= 64).
        no_epilogue / accumulate: a convolution whose K is split over two ops (the accumulator stays in TMEM in between, so the
        planes the first half read can be freed before the second half's inputs are produced)."""
        g = dict(self.lv[level])
        if n_mt is not None:               # a slice of the unit's rows: accumulator row 0 = unit row row_base (a multiple of 128)
            assert row_base % 128 == 0 and row_base + 128 * n_mt <= _ceil(g["R"], 128)
            g["n_mt"] = n_mt
        N = sum(cv.cout_pad for cv in cvs)
        assert N in (16, 64, 128), N
        if g["n_mt"] * N > 256:
            raise NotImplementedError("fused neck: accumulator does not fit the TMEM allocation")
        if no_epilogue:
            assert not outs and head < 0 and not pair and not deint
        elif deint:
            assert N <= 64 and len(outs) == 2 and g["W"] % 2 == 0 and not pair
        else:
            assert len(outs) == (N + 63) // 64
        kb_first = len(self.kbs)
        for src, shift, wb in kblocks:
            assert wb.shape == (N, 64), (wb.shape, N)
            self.kbs.append((int(src), int(shift), src_global >= 0))
            self.wblocks.append(wb)
        bias = torch.cat([cv.bias.float() for cv in cvs])
        assert bias.numel() == N
        flags = sum(1 << j for j, pr in enumerate(pair) if pr) | (256 if deint else 0) | (512 if no_epilogue else 0) | (1024 if accumulate else 0)
        if flags & 3:
            assert g["W"] % 2 == 0 and g["Wp"] % 2 == (0 if self.G > 1 else 1) and N >= 64, "pair-averaged outputs: even width"
        nxt = self.lv.get(level + 1)
        Wg = self.Ws[src_global] + 1 if src_global >= 0 else 0               # rows of one clip in the backbone's flat layout
        tpc = 0
        if src_global >= 0:
            tpc = g["n_mt"] if self.G == 1 else g["Wp"] // 128                # M tiles per clip (0: all clips of the unit in one tile)
        op = [CONV, g["n_mt"], N, kb_first, len(kblocks), g["R"], g["Wp"], g["W"], self.nbias, outs[0] if outs else -1,
              outs[1] if len(outs) > 1 else -1, head, flags, self.wrows, src_global, ACT_LRELU,
              1 if src_global >= 0 else max(1, 128 // N),                    # 16: K blocks per ring slot (smem-sourced convs)
              self.G, Wg, tpc,                                               # 17 clips per unit, 18 global rows per clip, 19 tiles per clip
              nxt["Wp"] if nxt else 0, row_base, 0, 0]                       # 20 clip pitch of the next level (pair / deint outputs),
        #                                                                      21 unit row of accumulator row 0
        for cv in cvs:
            assert cv.act == ACT_LRELU
        self.ops.append(op)
        self.biases.append(bias)
        self.nbias += N
        self.wrows += N * len(kblocks)

    def _taps3(self, cv, planes: List[int]):
        """3-tap (1 x 3 at H = 1: the middle row of the 3 x 3 filter) stride-1 convolution over the given planes (= torch.cat)."""
        assert cv.kh == 3 and cv.kw == 3 and cv.sh == 1 and cv.sw == 1 and cv.cin_pad == 64 * len(planes), (cv.name, cv.cin_pad, len(planes))
        return [(pl, kw - 1, self._blk([cv], 1, kw, c)) for c, pl in enumerate(planes) for kw in range(3)]

    def _taps1(self, cvs, planes: List[int]):
        for cv in cvs:
            assert cv.kh == 1 and cv.kw == 1 and cv.cin_pad == 64 * len(planes), (cv.name, cv.cin_pad, len(planes))
        return [(pl, 0, self._blk(cvs, 0, 0, c)) for c, pl in enumerate(planes)]

    def _taps_global(self, cvs, fi: int):
        """1 x 1 convolution(s) of the H-mean of backbone map ``fi``: K runs over the Hp * C channels of a flat column."""
        H, Cc = self.Hs[fi], self.Cs[fi]
        Hp = H + 1 if H > 1 else 1
        for cv in cvs:
            assert cv.kh == 1 and cv.kw == 1 and cv.cin_pad == Cc, (cv.name, cv.cin_pad, Cc)
        # the halo row h = H of a flat column is zero: its K blocks are left out (they were streamed against zero weights)
        return [(h * (Cc // 64) + c, 0, self._blk(cvs, 0, 0, c, 1.0 / H)) for h in range(H) for c in range(Cc // 64)]

    def _ew(self, typ: int, level_out: int, a: int, b: int, c: int = 0, d: int = 0, wp_in: int = 0):
        g = self.lv[level_out]
        self.ops.append([typ, a, b, c, d, g["R"], g["Wp"], g["W"], wp_in] + [0] * 15)

    def _dump(self, name: str, plane: int, level: int):
        if not self.debug:
            return
        rows = self.lv[level]["bytes"] // 128
        self.dumps[name] = (self.dump_elems, rows, level)
        self.ops.append([DUMP, plane, rows, self.dump_elems, 0] + [0] * 19)
        self.dump_elems += rows * 64

    def _build(self):
        n, rep, P, F = self.eng.n, self.eng.rep, self._plane, self._free
        lv = self.lv
        # ---- CSPSPPF (modules/_common.py:204-215) at level 4; conv1 and conv2 read the same map: one N = 128 convolution
        a1, y = P(4), P(4)
        self._conv([n["sp1"], n["sp2"]], 4, self._taps_global([n["sp1"], n["sp2"]], 3), [a1, y], src_global=3)
        self._dump("a1", a1, 4); self._dump("y", y, 4)
        a2 = P(4)
        self._conv([n["sp3"]], 4, self._taps3(n["sp3"], [a1]), [a2]); F(a1)
        self._dump("a2", a2, 4)
        a3 = P(4)
        self._conv([n["sp4"]], 4, self._taps1([n["sp4"]], [a2]), [a3]); F(a2)
        m1, m2, m3 = P(4), P(4), P(4)
        self._ew(POOLS, 4, a3, m1, m2, m3)
        self._dump("a3", a3, 4); self._dump("m1", m1, 4); self._dump("m3", m3, 4)
        b1 = P(4)
        self._conv([n["sp5"]], 4, self._taps1([n["sp5"]], [a3, m1, m2, m3]), [b1])
        self._dump("b1", b1, 4)
        for x in (a3, m1, m2, m3):
            F(x)
        b2 = P(4)
        self._conv([n["sp6"]], 4, self._taps3(n["sp6"], [b1]), [b2]); F(b1)
        p4 = [P(4), P(4)]
        self._conv([n["sp7"]], 4, self._taps1([n["sp7"]], [b2, y]), p4); F(b2); F(y)
        self._dump("p4_0", p4[0], 4); self._dump("p4_1", p4[1], 4)
        # ---- BiC3 (:179-185): cat[conv_c1(f3'), pairavg(conv_c0(f2')), up2(p4)] -> conv_out; f2' also feeds BiC2's conv_c1
        c1_3 = P(3)
        self._conv([n["b3c1"]], 3, self._taps_global([n["b3c1"]], 2), [c1_3], src_global=2)
        # conv_c0 of BiC3 (-> x0.5, written pair-averaged straight into the level-3 plane) and conv_c1 of BiC2 read the same map
        c0_3, c1_2 = P(3), P(2)
        self._conv([n["b3c0"], n["b2c1"]], 2, self._taps_global([n["b3c0"], n["b2c1"]], 1), [c0_3, c1_2], src_global=1, pair=(True, False))
        u3 = [P(3), P(3)]
        for i in range(2):
            self._ew(UP2, 3, p4[i], u3[i], wp_in=lv[4]["Wp"])
        # the output of a convolution may overwrite its own inputs (the epilogue starts after the last MMA has retired)
        bic3 = u3
        self._conv([n["b3o"]], 3, self._taps1([n["b3o"]], [c1_3, c0_3, u3[0], u3[1]]), bic3); F(c1_3); F(c0_3)
        self._dump("b3_0", bic3[0], 3)
        r1 = [P(3), P(3)]
        self._conv([rep["rep_block3_1"][0]["deploy"]], 3, self._taps3(rep["rep_block3_1"][0]["deploy"], bic3), r1); F(bic3[0]); F(bic3[1])
        p3 = [P(3), P(3)]
        self._conv([rep["rep_block3_1"][1]["deploy"]], 3, self._taps3(rep["rep_block3_1"][1]["deploy"], r1), p3); F(r1[0]); F(r1[1])
        self._dump("p3_0", p3[0], 3); self._dump("p3_1", p3[1], 3)
        # ---- BiC2
        c0_2 = P(2)
        if self.G > 1 and lv[1]["n_mt"] > 2:
            # the level-1 map of a unit spans more M tiles than the ring can hold A slots for: one pass per clip (the 8 weight
            # blocks of 8 KB are streamed again), each writing its clip's rows of the pair-averaged level-2 plane
            tiles = lv[1]["Wp"] // 128
            assert lv[1]["Wp"] % 128 == 0 and tiles <= 2
            for c in range(self.G):
                self._conv([n["b2c0"]], 1, self._taps_global([n["b2c0"]], 0), [c0_2], src_global=0, pair=(True,),
                           row_base=c * lv[1]["Wp"], n_mt=tiles)
        else:
            self._conv([n["b2c0"]], 1, self._taps_global([n["b2c0"]], 0), [c0_2], src_global=0, pair=(True,))
        # BiC2's conv_out (:184) with its K split over two ops: the conv_c1 / conv_c0 halves are accumulated (and their 2 level-2
        # planes freed) BEFORE the x2-upsampled halves are produced - the four input planes never coexist (the peak of the
        # shared-memory plan, 88 -> 56 KB per clip, is what lets two clips share a CTA pass)
        kb_b2o = self._taps1([n["b2o"]], [c1_2, c0_2, 0, 0])
        self._conv([n["b2o"]], 2, kb_b2o[:2], [], no_epilogue=True); F(c1_2); F(c0_2)
        u2 = [P(2), P(2)]
        for i in range(2):
            self._ew(UP2, 2, p3[i], u2[i], wp_in=lv[3]["Wp"])
        bic2 = u2
        kb_b2o = self._taps1([n["b2o"]], [0, 0, u2[0], u2[1]])
        self._conv([n["b2o"]], 2, kb_b2o[2:], bic2, accumulate=True)
        self._dump("b2_0", bic2[0], 2); self._dump("b2_1", bic2[1], 2)
        # ---- RepBlock2_1 -> n2 (sm head); its first conv writes over its own first input plane
        q1 = bic2[0]
        self._conv([rep["rep_block2_1"][0]["deploy"]], 2, self._taps3(rep["rep_block2_1"][0]["deploy"], bic2), [q1]); F(bic2[1])
        # ---- conv2_downsample (3 x 3, stride (1, 2), pad 1) on even / odd planes: out[k] = W0 odd[k-1] + W1 even[k] + W2 odd[k];
        #      the head conv in front of it writes its output split into those planes (no n2 plane, no DEINT op)
        e2, o2 = P(3), P(3)
        self._conv([rep["rep_block2_1"][1]["deploy"]], 2, self._taps3(rep["rep_block2_1"][1]["deploy"], [q1]), [e2, o2], head=0, deint=True); F(q1)
        d2 = [P(3), P(3)]
        self._conv([n["ds2"]], 3, self._taps_s2(n["ds2"], e2, o2), d2); F(e2); F(o2)
        q2 = P(3)
        self._conv([rep["rep_block3_2"][0]["deploy"]], 3, self._taps3(rep["rep_block3_2"][0]["deploy"], p3 + d2), [q2])
        for x in p3 + d2:
            F(x)
        e3, o3 = P(4), P(4)
        self._conv([rep["rep_block3_2"][1]["deploy"]], 3, self._taps3(rep["rep_block3_2"][1]["deploy"], [q2]), [e3, o3], head=1, deint=True); F(q2)
        d3 = [P(4), P(4)]
        self._conv([n["ds3"]], 4, self._taps_s2(n["ds3"], e3, o3), d3); F(e3); F(o3)
        q3 = P(4)
        self._conv([rep["rep_block4_1"][0]["deploy"]], 4, self._taps3(rep["rep_block4_1"][0]["deploy"], p4 + d3), [q3])
        for x in p4 + d3:
            F(x)
        n4 = P(4)
        self._conv([rep["rep_block4_1"][1]["deploy"]], 4, self._taps3(rep["rep_block4_1"][1]["deploy"], [q3]), [n4], head=2); F(q3); F(n4)
        assert not self.pool.live, self.pool.live

    def _taps_s2(self, cv, even: int, odd: int):
        assert cv.kh == 3 and cv.kw == 3 and (cv.sh, cv.sw) == (1, 2) and (cv.ph, cv.pw) == (1, 1) and cv.cin_pad == 64, cv.name
        return [(odd, -1, self._blk([cv], 1, 0, 0)), (even, 0, self._blk([cv], 1, 1, 0)), (odd, 0, self._blk([cv], 1, 2, 0))]

    def _finish(self):
        dev = self.dev
        offs, high = self.pool.place()
        res = lambda v: offs[v - TOKEN] if v >= TOKEN else v      # noqa: E731
        plane_fields = {CONV: (9, 10), POOLS: (1, 2, 3, 4), PAIRAVG: (1, 2), UP2: (1, 2), DEINT: (1, 2, 3), DUMP: (1,)}
        for op in self.ops:
            for f in plane_fields[op[0]]:
                op[f] = res(op[f])
        # smem-sourced K block: what the tap adds to the low descriptor word (plane offset + (guard row + shift) rows, in 16-byte
        # units); global-sourced: the 64-channel chunk of the input map
        self.kbs = [(src, 0) if glob else ((res(src) + (1 + sh) * 128) >> 4, 0) for src, sh, glob in self.kbs]
        self.pool_bytes = _ceil(high, 1024)
        tables = len(self.ops) * 96 + len(self.kbs) * 8 + 8 + 18 * 8 + 16          # (the biases stay in global memory)
        self.n_slots = min(8, (SMEM_MAX - 1024 - self.pool_bytes - tables) // SLOT)
        need = 1 + max(op[1] for op in self.ops if op[0] == CONV and op[14] >= 0)     # A tiles of one K block + its weight block
        if self.n_slots < need + 1:
            raise NotImplementedError(f"fused neck: shared-memory plan does not fit (pool {self.pool_bytes} B, {self.n_slots} ring slots)")
        self.wblob = torch.cat(self.wblocks, 0).to(dev, torch.bfloat16).contiguous()
        # the last block of the blob may be read with a larger box than its N: pad so that every box stays inside the array
        self.wblob = torch.cat([self.wblob, torch.zeros(128, 64, device=dev, dtype=torch.bfloat16)], 0).contiguous()
        assert self.wblob.shape[0] == self.wrows + 128
        self.bias = torch.cat(self.biases).to(dev, torch.float32).contiguous()
        self.kbs_t = torch.tensor(self.kbs, dtype=torch.int32, device=dev).contiguous()
        for op in self.ops:
            if op[0] == DUMP:
                op[4] = self.dump_elems          # per-clip stride of the debug buffer
        self.ops_t = torch.tensor(self.ops, dtype=torch.int32, device=dev).contiguous()
        assert self.ops_t.shape[1] == 24
        self.dbg = torch.zeros(0, dtype=torch.bfloat16, device=dev)

    # ------------------------------------------------------------------ launch
    def run(self, lib, fmaps: Sequence[torch.Tensor], heads: Sequence[torch.Tensor], stream, dbg: Optional[torch.Tensor] = None):
        """fmaps: the four flat backbone maps [B, Wp, Hp, C] bf16; heads: three fp32 tensors [B, 1, W, ld]."""
        B = fmaps[0].shape[0]
        fk, fr, fb = [], [], []
        for i, f in enumerate(fmaps):
            _, Wp, Hp, Cc = f.shape
            assert f.is_contiguous() and f.dtype == torch.bfloat16 and Wp == self.Ws[i] + 1 and Cc == self.Cs[i], (i, tuple(f.shape))
            assert Hp == (self.Hs[i] + 1 if self.Hs[i] > 1 else 1)
            fk.append(Hp * Cc)
            fr.append(Wp)
            # rows of one TMA box of map i: a whole clip (rounded up to 8 rows) when it fits an M tile, else 128; with several clips
            # per unit the clip pitch (a power of two >= W + 1)
            P = self.lv[i + 1]["Wp"]
            fb.append(min(128, P if self.G > 1 else _ceil(Wp, 8)))
        ld = heads[0].shape[-1]
        for i, h in enumerate(heads):
            assert h.dtype == torch.float32 and h.is_contiguous() and h.shape[-1] == ld and h.shape[-2] == self.Ws[i + 1] and h.shape[0] == B
        fm = (C.c_void_p * 4)(*[f.data_ptr() for f in fmaps])
        hd = (C.c_void_p * 3)(*[h.data_ptr() for h in heads])
        rc = lib.yad_neck_fused(fm, (C.c_int32 * 4)(*fk), (C.c_int32 * 4)(*fr), (C.c_int32 * 4)(*fb), B, self.G, self.wblob.data_ptr(), self.wblob.shape[0],
                                self.bias.data_ptr(), self.nbias, self.ops_t.data_ptr(), len(self.ops), self.kbs_t.data_ptr(), len(self.kbs),
                                self.pool_bytes, self.n_slots, hd, (C.c_int32 * 3)(*self.Ws[1:]), ld,
                                0 if dbg is None else dbg.data_ptr(), stream)
        _lib.check(rc, "neck_fused")
