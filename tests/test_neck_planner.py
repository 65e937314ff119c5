"""Host-side compiler of the fused neck (yad_b200/neck_fused.py; reference: MultiScaleFmapModule.forward, modules/_common.py:241-265):
the shared-memory plan, the program's well-formedness and the fall-back from two clips per pass to one - on a mock engine, no GPU."""
import types

import pytest
import torch

import yad_b200  # noqa: F401
from yad_b200 import neck_fused as nf


class _CV:
    def __init__(self, name, cin, cout, k, sw=1):
        self.name = name
        self.cin_pad = (cin + 63) // 64 * 64
        self.cout_pad = 16 if cout == 15 else cout
        self.kh = self.kw = k
        self.sh, self.sw = 1, sw
        self.ph = self.pw = k // 2
        self.act = nf.ACT_LRELU
        self.w = torch.zeros(self.cout_pad, k * k * self.cin_pad)
        self.bias = torch.zeros(self.cout_pad)


def _engine():
    n = {"sp1": _CV("sp1", 512, 64, 1), "sp2": _CV("sp2", 512, 64, 1), "sp3": _CV("sp3", 64, 64, 3), "sp4": _CV("sp4", 64, 64, 1),
         "sp5": _CV("sp5", 256, 64, 1), "sp6": _CV("sp6", 64, 64, 3), "sp7": _CV("sp7", 128, 128, 1), "b3c1": _CV("b3c1", 256, 64, 1),
         "b3c0": _CV("b3c0", 128, 64, 1), "b2c1": _CV("b2c1", 128, 64, 1), "b3o": _CV("b3o", 256, 128, 1), "b2c0": _CV("b2c0", 64, 64, 1),
         "b2o": _CV("b2o", 256, 128, 1), "ds2": _CV("ds2", 15, 128, 3, 2), "ds3": _CV("ds3", 15, 128, 3, 2)}
    rep = {"rep_block3_1": [{"deploy": _CV("r31a", 128, 128, 3)}, {"deploy": _CV("r31b", 128, 128, 3)}],
           "rep_block2_1": [{"deploy": _CV("r21a", 128, 64, 3)}, {"deploy": _CV("r21b", 64, 15, 3)}],
           "rep_block3_2": [{"deploy": _CV("r32a", 256, 15, 3)}, {"deploy": _CV("r32b", 15, 15, 3)}],
           "rep_block4_1": [{"deploy": _CV("r41a", 256, 15, 3)}, {"deploy": _CV("r41b", 15, 15, 3)}]}
    return types.SimpleNamespace(dev=torch.device("cpu"), n=n, rep=rep)


def _neck(W1, G=None):
    return nf.FusedNeck(_engine(), [8, 4, 2, 1], [W1, W1 // 2, W1 // 4, W1 // 8], [64, 128, 256, 512], G=G)


@pytest.mark.parametrize("W1", [8, 40, 88, 240])
@pytest.mark.parametrize("G", [1, 2])
def test_plan_fits_and_is_well_formed(W1, G):
    f = _neck(W1, G)
    assert f.G == G
    tables = len(f.ops) * 96 + len(f.kbs) * 8 + 8 + 18 * 8 + 16
    assert 1024 + f.pool_bytes + f.n_slots * nf.SLOT + tables <= nf.SMEM_MAX
    need = 1 + max(op[1] for op in f.ops if op[0] == nf.CONV and op[14] >= 0)
    assert 3 <= need + 1 <= f.n_slots <= 8
    lv = f.lv
    if G == 2:       # power-of-two pitches that nest, at least one swizzle atom, wide enough for the clip and its halo cell
        assert all(lv[i]["Wp"] == 2 * lv[i + 1]["Wp"] for i in (1, 2, 3)) and lv[4]["Wp"] >= 8 and lv[4]["Wp"] & (lv[4]["Wp"] - 1) == 0
    assert all(lv[i]["Wp"] >= lv[i]["W"] + 1 for i in lv)
    n_split = 0
    for op in f.ops:
        assert len(op) == 24
        if op[0] != nf.CONV:
            continue
        n_mt, N, nkb, R, P = op[1], op[2], op[4], op[5], op[6]
        assert N in (16, 64, 128) and n_mt * N <= 256 and 1 <= n_mt <= 4 and nkb >= 1
        assert op[17] == G and R == G * P and op[21] % 128 == 0 and op[21] + 128 * n_mt <= (R + 127) // 128 * 128
        assert 1 <= op[16] <= 8 and op[16] * N <= 128
        flags = op[12]
        if flags & 512:                       # first half of a split-K convolution: no outputs, the next conv accumulates
            assert op[9] == -1 and op[10] == -1 and op[11] == -1
            n_split += 1
        if flags & 256:                       # de-interleaved output: two planes, a pitch for them
            assert op[9] >= 0 and op[10] >= 0 and op[20] * 2 == P + (1 if G == 1 else 0) or op[20] == (op[7] >> 1) + 1
        if op[14] >= 0:
            assert op[18] == f.Ws[op[14]] + 1 and (op[19] == 0) == (G > 1 and P < 128)
    assert n_split == 1 and sum(1 for op in f.ops if op[0] == nf.CONV and op[12] & 1024) == 1
    # every plane offset is 1024-aligned and inside the pool
    for op in f.ops:
        fields = {nf.CONV: (9, 10), nf.POOLS: (1, 2, 3, 4), nf.UP2: (1, 2)}.get(op[0], ())
        for i in fields:
            assert op[i] == -1 or (op[i] % 1024 == 0 and 0 <= op[i] < f.pool_bytes)


def test_two_clip_plan_of_a_60s_clip_reaches_its_liveness_peak():
    f = _neck(240)
    assert f.G == 2 and f.n_slots >= 5
    planes = f.pool.planes
    T = max(d for _, _, d in planes)
    peak = max(sum(n for n, b, d in planes if b <= t < d) for t in range(T + 1))
    assert f.pool_bytes == peak * 1024          # the multi-order first-fit packer wastes nothing here (120 KB)
    # the level-1 conv runs once per clip: two ops on map 0, accumulator rows [0, 256) and [256, 512)
    l1 = [op for op in f.ops if op[0] == nf.CONV and op[14] == 0]
    assert [op[21] for op in l1] == [0, 256] and all(op[1] == 2 for op in l1)


def test_falls_back_to_one_clip_per_pass_when_two_do_not_fit():
    f = _neck(480)                              # 120 s clips: the two-clip plan needs more than 4 M tiles at level 1
    assert f.G == 1
    with pytest.raises(NotImplementedError):
        _neck(480, G=2)
    with pytest.raises(ValueError):
        nf.FusedNeck(_engine(), [8, 4, 2, 1], [240, 120, 60, 31], [64, 128, 256, 512])
