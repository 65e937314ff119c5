// The whole RepBi-PAN neck of the default (ResNet) net as ONE persistent tcgen05 kernel.
//
// Replaces MultiScaleFmapModule.forward (modules/_common.py:241-265) in deploy form at H = 1: the H-mean of the four backbone
// maps (:248-252), CSPSPPF (:204-215), the two BiC blocks (:179-185), four RepBlocks (:148-158), the two stride-(1,2)
// downsample convs (:238-239) and the permute to [B, G, 15] (:259-264) - 21 convolutions, 3 cascaded max-pools, 2 bilinear x2,
// 2 bilinear x0.5 and the three H-means, which the layer-by-layer path runs as 31 launches of 10-27 us each (488 us per
// 512-clip step for 1.9 % of the network's FLOPs: every launch is latency, none is throughput).
//
// Design: one CTA walks a host-compiled PROGRAM (a list of ops, see neck_fused.py) for one clip at a time; every intermediate
// activation of the clip lives in shared memory in the tensor core's own operand layout, so a convolution's A operand is a
// descriptor into the previous convolution's output:
//   * activation plane = 64 channels x rows, 128-byte rows, 128B-swizzled in 8-row atoms (the canonical K-major SWIZZLE_128B
//     UMMA layout).  Row s holds logical row r = s - 1 (row 0 is a zero guard row); logical rows are the flat halo layout of
//     conv_flat.cu at H = 1: r = clip * (W + 1) + w, the cell w = W of every clip is a zero halo.  A 3-tap convolution is then
//     three MMAs per 64-channel chunk whose A descriptors start at rows s - 1, s, s + 1 (any 128-byte row is a legal start, see
//     conv_flat.cu); a torch.cat is a list of planes; the stride-2 downsample convs read de-interleaved even / odd planes.
//   * the backbone maps are read straight out of their flat [B, Wp, Hp, C] layout by TMA: the column (b, w) is Hp * C
//     contiguous channels, so "H-mean, then 1x1 conv" is one GEMM with K = Hp * C against the 1x1 weights replicated Hp times
//     and scaled by 1/H (exact in bf16: H is a power of two; the halo row h = H is zero) - no H-mean kernel, no pooled tensor.
//   * weights stream from L2 through one ring of 16 KB slots ([N x 64] blocks, one per K block) shared with the A tiles of the
//     global-sourced convs; the producer warp runs ahead of the MMAs across op boundaries (weights do not depend on data).
//   * accumulators live in TMEM ([128 x N] per M tile); the 8 epilogue warps add the bias, apply LeakyReLU(0.2), zero the halo
//     rows and write bf16 planes (and the fp32 head rows to global memory); the same warps run the element-wise ops (pools,
//     bilinear resizes, de-interleave) between convolutions.
// Ops are strictly sequential per clip (the neck is a dependency chain); clips are independent, one CTA per SM loops over them.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace yad {

constexpr int NK_THREADS = 320;   // warp 0: TMA producer, warp 1: MMA issuer, warps 2..9: epilogue + element-wise ops
constexpr int NK_EPI = 256;       // two warps per TMEM lane quadrant, each taking every second 32-column block
constexpr int NK_SLOT = 16384;    // ring slot: 128 rows x 128 B
constexpr int NK_MAX_SLOTS = 8;
constexpr int NK_MAX_MT = 4;
constexpr int NK_TMEM_COLS = 256;
constexpr int NK_TL = 8;          // timeline stamps per op (yad_neck_fused_set_timeline)
constexpr int NK_MAX_KG = 8;      // K blocks per ring slot (N = 16: 8 x 2 KB)

enum { NK_CONV = 0, NK_POOLS = 1, NK_PAIRAVG = 2, NK_UP2 = 3, NK_DEINT = 4, NK_DUMP = 5 };

// One op = 16 int32 (the meaning of v[1..] depends on the type; byte offsets are relative to the 1024-aligned smem base).
//   CONV   : 1 n_mt, 2 N, 3 kb_first, 4 kb_count, 5 R, 6 Wp, 7 W, 8 bias_off, 9 out_plane0, 10 out_plane1, 11 head (-1: none),
//            12 flags (bit j: output plane j is written PAIR-AVERAGED, out[k] = (y[2k] + y[2k+1]) / 2 of the bf16-rounded rows,
//            at the next level's geometry W / 2 - the bilinear x0.5 of BiC's conv_c0 branch folded into the epilogue),
//            (bit 8: DE-INTERLEAVED output for the stride-2 conv that follows: row w of clip c goes to row c * (W / 2 + 1) + w / 2
//            of plane out_plane0 (w even) or out_plane1 (w odd) - the even / odd split folded into the epilogue),
//            13 wrow, 14 src_global (-1: planes in smem, else input map 0..3), 15 act,
//            (bit 9: NO EPILOGUE - the first half of a convolution whose K is split over two ops; bit 10: ACCUMULATE onto what
//            such an op left in TMEM),
//            16 kg = K blocks per weight-ring slot (kg * N <= 128 rows), 17 G = clips per unit, 18 rows of one clip in the
//            backbone's flat layout (global-sourced convs), 19 M tiles per clip (0: all clips of the unit share one tile),
//            20 clip pitch of the next level (pair-averaged / de-interleaved outputs), 21 unit row of accumulator row 0
//            (a convolution may cover a 128-aligned slice of the unit's rows).
//            Rows: a UNIT = G clips processed together by one CTA pass; clip c of the unit owns rows [c * P, c * P + P) of every
//            plane of a level (P = field 6: W + 1 for G = 1, a power of two >= W + 1 for G = 2), its cells w >= W are zero.
//   POOLS  : 1 in, 2 out1, 3 out2, 4 out3, 5 R, 6 Wp, 7 W
//   PAIRAVG: 1 in, 2 out, 5 R_out, 6 Wp_out, 7 W_out, 8 Wp_in            (out[k] = (in[2k] + in[2k+1]) / 2)
//   UP2    : 1 in, 2 out, 5 R_out, 6 Wp_out, 7 W_out, 8 Wp_in            (bilinear x2, align_corners = False)
//   DEINT  : 1 in, 2 out_even, 3 out_odd, 5 R_out, 6 Wp_out, 7 W_out, 8 Wp_in
//   DUMP   : 1 plane, 2 rows, 3 destination offset (bf16 elements) in the debug buffer, 4 per-clip stride (elements)
constexpr int NK_OP_WORDS = 24;
struct NkOp {
  int32_t v[NK_OP_WORDS];
};
struct NkKb {
  int32_t src;     // smem-sourced conv: (plane byte offset + (1 + tap shift) * 128) >> 4 = what the tap adds to the low descriptor
                   // word; global-sourced conv: 64-channel chunk index of the input map
  int32_t pad;
};

struct NkParams {
  long long* tlog;     // debug timeline (yad_neck_fused_set_timeline): CTA 0 records clock64 per op, NK_TL stamps each; NULL in production
  int32_t n_ops, n_kb, n_clips, n_slots;
  int32_t G, n_units;  // clips per unit, units = ceil(n_clips / G)
  int32_t pool_bytes, n_bias;
  int32_t head_W[3];
  int32_t head_ld;
  int32_t a_bytes[4];  // bytes of one A box of input map i (rows of one clip, rounded up to 8, at most 128, x 128 B)
  int32_t tl_iter;     // timeline: which of CTA 0's clips is recorded (0 = first: cold, nothing prefetched; 1 = steady state)
};

__device__ __forceinline__ void nk_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void nk_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// K-major SWIZZLE_128B descriptor (see conv_flat.cu): SBO = 1024 B, version 1, layout SWIZZLE_128B, base_offset 0
constexpr uint32_t NK_DESC_HI = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint64_t nk_desc(uint32_t smem_addr) {
  return ((uint64_t)NK_DESC_HI << 32) | (uint64_t)(((smem_addr & 0x3FFFFu) >> 4) | (1u << 16));
}

// the same from a precomputed low word ((address >> 4) | 1 << 16): adding 8 per row, 2 per 16 K elements needs no re-masking
__device__ __forceinline__ uint64_t nk_desc64(uint32_t lo) { return ((uint64_t)NK_DESC_HI << 32) | (uint64_t)lo; }

// element-wise maximum of 8 packed bf16 values
__device__ __forceinline__ uint4 nk_max_bf16x8(const uint4 a, const uint4 b) {
  const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
  uint32_t rw[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const __nv_bfloat162 m = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&aw[e]), *reinterpret_cast<const __nv_bfloat162*>(&bw[e]));
    rw[e] = *reinterpret_cast<const uint32_t*>(&m);
  }
  return make_uint4(rw[0], rw[1], rw[2], rw[3]);
}
// 0.75 a + 0.25 b on 8 packed bf16 values, rounded once
__device__ __forceinline__ uint4 nk_lerp_bf16x8(const uint4 a, const uint4 b) {
  const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
  uint32_t rw[4];
  const __nv_bfloat162 c75 = __floats2bfloat162_rn(0.75f, 0.75f), c25 = __floats2bfloat162_rn(0.25f, 0.25f);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const __nv_bfloat162 q = __hmul2(c25, *reinterpret_cast<const __nv_bfloat162*>(&bw[e]));      // exact
    const __nv_bfloat162 m = __hfma2(c75, *reinterpret_cast<const __nv_bfloat162*>(&aw[e]), q);
    rw[e] = *reinterpret_cast<const uint32_t*>(&m);
  }
  return make_uint4(rw[0], rw[1], rw[2], rw[3]);
}

// 16-byte chunk j (8 channels) of row s of a plane
__device__ __forceinline__ uint4* nk_chunk(uint8_t* plane, int s, int j) {
  return reinterpret_cast<uint4*>(plane + s * 128 + ((j ^ (s & 7)) << 4));
}
__device__ __forceinline__ void nk_unpack(const uint4 u, float* f) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    f[2 * e] = __uint_as_float(w[e] << 16);
    f[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 nk_pack(const float* f) {
  uint32_t o[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
    o[e] = *reinterpret_cast<uint32_t*>(&h2);
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}

__global__ void __launch_bounds__(NK_THREADS, 1)
neck_fused_kernel(const __grid_constant__ CUtensorMap map_in0, const __grid_constant__ CUtensorMap map_in1,
                  const __grid_constant__ CUtensorMap map_in2, const __grid_constant__ CUtensorMap map_in3,
                  const __grid_constant__ CUtensorMap map_w16, const __grid_constant__ CUtensorMap map_w64,
                  const __grid_constant__ CUtensorMap map_w128, const NkParams p, const NkOp* __restrict__ g_ops,
                  const NkKb* __restrict__ g_kbs, const float* __restrict__ g_bias, float* __restrict__ head0,
                  float* __restrict__ head1, float* __restrict__ head2, __nv_bfloat16* __restrict__ dbg) {
  extern __shared__ __align__(1024) uint8_t nk_smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(nk_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* ring = base + p.pool_bytes;
  NkOp* s_ops = reinterpret_cast<NkOp*>(ring + (size_t)p.n_slots * NK_SLOT);
  NkKb* s_kbs = reinterpret_cast<NkKb*>(s_ops + p.n_ops);
  // (the biases stay in global memory: 7.6 KB of shared memory buy a sixth ring slot on the two-clip plan; the epilogue's
  // broadcast float4 loads hit L1 after a CTA's first unit)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(s_kbs + p.n_kb) + 7) & ~(uintptr_t)7);
  uint64_t* empty_bar = full_bar + NK_MAX_SLOTS;
  uint64_t* acc_full = empty_bar + NK_MAX_SLOTS;
  uint64_t* op_done = acc_full + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(op_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // the program, the K-block list and the biases are launch constants (written once at pack time): loaded before pdl_wait
  for (int i = threadIdx.x; i < p.n_ops * NK_OP_WORDS; i += NK_THREADS) reinterpret_cast<int32_t*>(s_ops)[i] = reinterpret_cast<const int32_t*>(g_ops)[i];
  for (int i = threadIdx.x; i < p.n_kb * 2; i += NK_THREADS) reinterpret_cast<int32_t*>(s_kbs)[i] = reinterpret_cast<const int32_t*>(g_kbs)[i];
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_in0);
    prefetch_tmap(&map_in1);
    prefetch_tmap(&map_in2);
    prefetch_tmap(&map_in3);
    prefetch_tmap(&map_w16);
    prefetch_tmap(&map_w64);
    prefetch_tmap(&map_w128);
    for (int s = 0; s < NK_MAX_SLOTS; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(op_done, NK_EPI / 32);       // one arrival per epilogue warp
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, NK_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();          // programmatic dependent launch: the backbone's last kernel is complete from here on
  pdl_trigger();

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      uint32_t slot = 0, phase = 0;
      for (int unit = blockIdx.x; unit < p.n_units; unit += gridDim.x) {
        for (int oi = 0; oi < p.n_ops; ++oi) {
          const NkOp& op = s_ops[oi];
          if (op.v[0] != NK_CONV) continue;
          const int n_mt = op.v[1], N = op.v[2], kb0 = op.v[3], nkb = op.v[4], srcg = op.v[14], kg = op.v[16];
          const CUtensorMap* mw = N == 16 ? &map_w16 : (N == 64 ? &map_w64 : &map_w128);
          if (srcg >= 0) {
            const CUtensorMap* ma = srcg == 0 ? &map_in0 : (srcg == 1 ? &map_in1 : (srcg == 2 ? &map_in2 : &map_in3));
            const uint32_t a_bytes = (uint32_t)p.a_bytes[srcg];
            const int P = op.v[6], G = op.v[17], Wg = op.v[18], tpc = op.v[19], t0 = op.v[21] >> 7;
            for (int k = 0; k < nkb; ++k) {
              const int c0 = s_kbs[kb0 + k].src * 64;
              for (int mt = 0; mt < n_mt; ++mt) {
                mbar_wait(&empty_bar[slot], phase ^ 1);
                uint8_t* dst = ring + (size_t)slot * NK_SLOT;
                if (tpc > 0) {          // this tile lies inside one clip: one box (rows past the clip land on its zeroed cells)
                  const int c = (t0 + mt) / tpc, w0 = (t0 + mt - c * tpc) << 7;
                  mbar_expect_tx(&full_bar[slot], a_bytes);
                  tma_load_2d(ma, &full_bar[slot], dst, c0, (unit * G + c) * Wg + w0);
                } else {                // all clips of the unit in this tile: one box of P rows per clip at row c * P
                  mbar_expect_tx(&full_bar[slot], a_bytes * (uint32_t)G);
                  for (int c = 0; c < G; ++c) tma_load_2d(ma, &full_bar[slot], dst + (size_t)(c * P) * 128, c0, (unit * G + c) * Wg);
                }
                if (++slot == (uint32_t)p.n_slots) { slot = 0; phase ^= 1; }
              }
              mbar_wait(&empty_bar[slot], phase ^ 1);
              mbar_expect_tx(&full_bar[slot], (uint32_t)N * 128u);
              tma_load_2d(mw, &full_bar[slot], ring + (size_t)slot * NK_SLOT, 0, op.v[13] + k * N);
              if (++slot == (uint32_t)p.n_slots) { slot = 0; phase ^= 1; }
            }
          } else {
            // smem-sourced conv: kg consecutive weight blocks share one ring slot (one barrier)
            for (int k0 = 0; k0 < nkb; k0 += kg) {
              const int cnt = nkb - k0 < kg ? nkb - k0 : kg;
              mbar_wait(&empty_bar[slot], phase ^ 1);
              mbar_expect_tx(&full_bar[slot], (uint32_t)(cnt * N) * 128u);
              for (int j = 0; j < cnt; ++j)
                tma_load_2d(mw, &full_bar[slot], ring + (size_t)slot * NK_SLOT + (size_t)(j * N) * 128, 0, op.v[13] + (k0 + j) * N);
              if (++slot == (uint32_t)p.n_slots) { slot = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (whole warp walks, one lane issues)
    // The tensor pipe queues only ONE MMA behind the executing one (conv_flat.cu) and a lone thread retires an instruction
    // every 5 - 10 cycles, so whatever this warp does between two MMAs - barrier poll, fence, election, table load, descriptor
    // moves, commit - is idle pipe time; the first version of this kernel (descriptors built with shift / mask / or per MMA,
    // one visit per K block) paid ~580 cycles per K block for 160 - 256 cycles of MMA.  Now (a) descriptor low words are sums
    // of precomputed terms (the host table carries the tap's share, see NkKb) and (b) kg K blocks share one ring slot = one
    // visit (poll, fence, election, commit) per 4 * kg MMAs: 280 (N = 16) / 480 (N = 64) / 600 (N = 128) cycles per K block.
    // Measured in round 2 and dropped (profiles/r02_neck_timeline_*.txt): a second issuing warp taking alternate slots into a
    // second accumulator set (the per-MMA issue cost, not the visit, dominates; the epilogue pays for the extra TMEM loads);
    // releasing the ring slots of a whole convolution in one burst of commits; visits of up to 4 slots / 8 K blocks, rolled
    // or unrolled (more bookkeeping per MMA than they save per visit); one polling lane per waiting epilogue warp; L2
    // prefetch of the next clip's maps by the producer (a prefetch occupies the TMA unit as long as the load it saves).
    uint32_t slot = 0, phase = 0, done_phase = 0;
    bool first_conv = true;
    const uint32_t plane_lo = (smem_u32(base) >> 4) | (1u << 16);       // + NkKb.src + 1024 * mt + 2 * kk
    const uint32_t ring_lo = (smem_u32(ring) >> 4) | (1u << 16);        // + 1024 * slot + 8 * N * j + 2 * kk
    const uint32_t n_slots = (uint32_t)p.n_slots;
    int iter = 0;
    for (int unit = blockIdx.x; unit < p.n_units; unit += gridDim.x, ++iter) {
      for (int oi = 0; oi < p.n_ops; ++oi) {
        const NkOp& op = s_ops[oi];
        if (op.v[0] == NK_CONV) {
          // Everything before this convolution is complete: the epilogue warps arrive on op_done at the end of the op
          // that PRECEDES a convolution (and of the program's last op), i.e. once per convolution - planes written and fenced
          // for the async proxy, accumulator drained.  One phase per convolution: the epilogue cannot get a phase ahead,
          // because it next waits for this convolution's accumulator.
          if (!first_conv) {
            mbar_wait(op_done, done_phase);
            done_phase ^= 1;
            tc_fence_after();
          }
          first_conv = false;
          const bool tl = p.tlog != nullptr && blockIdx.x == 0 && lane == 0 && iter == p.tl_iter;
          if (tl) p.tlog[oi * NK_TL + 0] = clock64();
          const int n_mt = op.v[1], N = op.v[2], kb0 = op.v[3], nkb = op.v[4], srcg = op.v[14], kg = op.v[16];
          const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
          uint32_t acc = (op.v[12] & 1024) ? 1u : 0u;      // second half of a split-K convolution: onto the first half's sums
          if (srcg >= 0) {
            for (int k = 0; k < nkb; ++k) {
              uint32_t a_slot[NK_MAX_MT];
#pragma unroll
              for (int mt = 0; mt < NK_MAX_MT; ++mt) {
                if (mt < n_mt) {
                  mbar_wait(&full_bar[slot], phase);
                  a_slot[mt] = slot;
                  if (++slot == n_slots) { slot = 0; phase ^= 1; }
                }
              }
              mbar_wait(&full_bar[slot], phase);
              const uint32_t w_slot = slot, b_lo = ring_lo + slot * (NK_SLOT >> 4);
              if (++slot == n_slots) { slot = 0; phase ^= 1; }
              tc_fence_after();
              if (elect_one()) {
#pragma unroll
                for (int mt = 0; mt < NK_MAX_MT; ++mt) {
                  if (mt < n_mt) {
                    const uint32_t a_lo = ring_lo + a_slot[mt] * (NK_SLOT >> 4);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                      umma_bf16(tmem_base + (uint32_t)(mt * N), nk_desc64(a_lo + 2 * kk), nk_desc64(b_lo + 2 * kk), idesc,
                                kk > 0 ? 1u : acc);
                  }
                }
#pragma unroll
                for (int mt = 0; mt < NK_MAX_MT; ++mt)
                  if (mt < n_mt) umma_commit(&empty_bar[a_slot[mt]]);
                umma_commit(&empty_bar[w_slot]);
                if (k == nkb - 1) umma_commit(acc_full);
              }
              acc = 1u;
              __syncwarp();
            }
          } else {
            const uint32_t n_units = (uint32_t)N * 8u;                  // one weight block = N rows x 128 B, in 16-byte units
            for (int k0 = 0; k0 < nkb; k0 += kg) {
              const int cnt = nkb - k0 < kg ? nkb - k0 : kg;
              mbar_wait(&full_bar[slot], phase);
              tc_fence_after();
              if (elect_one()) {
                uint32_t b_lo = ring_lo + slot * (NK_SLOT >> 4);
                for (int j = 0; j < cnt; ++j) {
                  uint32_t a_lo = plane_lo + (uint32_t)s_kbs[kb0 + k0 + j].src;
                  for (int mt = 0; mt < n_mt; ++mt) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                      umma_bf16(tmem_base + (uint32_t)(mt * N), nk_desc64(a_lo + 2 * kk), nk_desc64(b_lo + 2 * kk), idesc,
                                kk > 0 ? 1u : acc);
                    a_lo += 1024u;                                      // next M tile: 128 rows x 128 B
                  }
                  acc = 1u;
                  b_lo += n_units;
                }
                umma_commit(&empty_bar[slot]);
                if (k0 + cnt == nkb) umma_commit(acc_full);
              }
              acc = 1u;
              __syncwarp();
              if (++slot == n_slots) { slot = 0; phase ^= 1; }
            }
          }
          if (tl) p.tlog[oi * NK_TL + 1] = clock64();
        }
      }
    }
  } else {
    // ===================================================================== epilogue + element-wise ops (256 threads)
    const int q = warp & 3;               // TMEM lane quadrant this warp may read
    const int half = (warp - 2) >> 2;     // which 32-column blocks of an accumulator this warp converts (b & 1 == half)
    const int te = threadIdx.x - 64;      // 0..255
    uint32_t acc_phase = 0;
    int iter = 0;
    for (int unit = blockIdx.x; unit < p.n_units; unit += gridDim.x, ++iter) {
      const bool tle = p.tlog != nullptr && blockIdx.x == 0 && iter == p.tl_iter;
      for (int oi = 0; oi < p.n_ops; ++oi) {
        const NkOp& op = s_ops[oi];
        nk_bar_sync(1, NK_EPI);           // the previous op's planes are complete (and no thread still reads what this op overwrites)
        const int type = op.v[0];
        if (type == NK_CONV) {
          const int n_mt = op.v[1], N = op.v[2], R = op.v[5], Wp = op.v[6], W = op.v[7], head = op.v[11], act = op.v[15];
          const float* bias = g_bias + op.v[8];
          // act(x) = max(x, slope * x): LeakyReLU(0.2), ReLU (slope 0) and identity (slope 1) without a branch per element
          const float slope = act == YAD_ACT_LRELU02 ? 0.2f : (act == YAD_ACT_RELU ? 0.0f : 1.0f);
          mbar_wait(acc_full, acc_phase);
          acc_phase ^= 1;
          tc_fence_after();
          if (tle && te == 0) p.tlog[oi * NK_TL + 2] = clock64();
          const int nb = (N + 31) >> 5;   // 32-column blocks (N = 16: one block, upper 16 columns unused)
          const int n_mt_e = (op.v[12] & 512) ? 0 : n_mt;        // first half of a split-K convolution: the sums stay in TMEM
          for (int mt = 0; mt < n_mt_e; ++mt) {
            const int r = op.v[21] + 128 * mt + q * 32 + lane, s = r + 1;
            const int c = r / Wp, w = r - c * Wp;
            const bool valid = r < R && w < W;
            const bool in_plane = r < R;  // rows past the level's R rows are not part of the plane (the next plane starts there)
            for (int b = half; b < nb; b += 2) {
              uint32_t vv[32];
              tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * N + 32 * b), vv);
              tmem_ld_wait();
              float x[32];
              const float4* b4 = reinterpret_cast<const float4*>(bias + 32 * b);
#pragma unroll
              for (int i4 = 0; i4 < 8; ++i4) {
                if (N == 16 && i4 >= 4) {
                  x[4 * i4] = x[4 * i4 + 1] = x[4 * i4 + 2] = x[4 * i4 + 3] = 0.0f;
                } else {
                  const float4 bb = __ldg(b4 + i4);
                  const float t0 = __uint_as_float(vv[4 * i4]) + bb.x, t1 = __uint_as_float(vv[4 * i4 + 1]) + bb.y;
                  const float t2 = __uint_as_float(vv[4 * i4 + 2]) + bb.z, t3 = __uint_as_float(vv[4 * i4 + 3]) + bb.w;
                  x[4 * i4] = fmaxf(t0, slope * t0);
                  x[4 * i4 + 1] = fmaxf(t1, slope * t1);
                  x[4 * i4 + 2] = fmaxf(t2, slope * t2);
                  x[4 * i4 + 3] = fmaxf(t3, slope * t3);
                }
              }
              if (head >= 0 && valid && unit * p.G + c < p.n_clips) {      // fp32 head rows [B, W, head_ld] for the decoder (N = 16)
                float* hp = (head == 0 ? head0 : (head == 1 ? head1 : head2)) +
                            ((int64_t)(unit * p.G + c) * p.head_W[head] + w) * p.head_ld;
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4)
                  if (4 * i4 < p.head_ld) reinterpret_cast<float4*>(hp)[i4] = make_float4(x[4 * i4], x[4 * i4 + 1], x[4 * i4 + 2], x[4 * i4 + 3]);
              }
              const int pj = op.v[9 + (b >> 1)];
              const int hb = 4 * (b & 1);
              if ((op.v[12] >> (b >> 1)) & 1) {
                // pair-averaged output (one clip per pass: r = w): rows (2k, 2k + 1) sit in adjacent lanes; both are rounded to
                // bf16 first (what the separate x0.5 pass read), averaged in fp32, and lane 2k writes row k of the half-width plane
                uint8_t* plane = base + pj;
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {
                  const uint4 mine = nk_pack(x + 8 * i4);
                  uint4 oth;
                  oth.x = __shfl_xor_sync(0xffffffffu, mine.x, 1);
                  oth.y = __shfl_xor_sync(0xffffffffu, mine.y, 1);
                  oth.z = __shfl_xor_sync(0xffffffffu, mine.z, 1);
                  oth.w = __shfl_xor_sync(0xffffffffu, mine.w, 1);
                  float fa[8], fb[8], av[8];
                  nk_unpack(mine, fa);
                  nk_unpack(oth, fb);
#pragma unroll
                  for (int e = 0; e < 8; ++e) av[e] = 0.5f * fa[e] + 0.5f * fb[e];
                  uint4 pk = nk_pack(av);
                  if (!valid) pk = make_uint4(0u, 0u, 0u, 0u);          // w = W (even): the halo cell of the half-width plane
                  if (in_plane && !(r & 1)) *nk_chunk(plane, (r >> 1) + 1, hb + i4) = pk;
                }
              } else if (op.v[12] & 256) {
                // de-interleaved output (N <= 64: one 64-channel block) in front of a stride-2 conv: even columns to out_plane0,
                // odd ones to out_plane1, both at the next level's geometry W / 2 (+ halo); the halo cell w = W (W even) lands
                // on the halo cell of the even plane, the same thread zeroes the one of the odd plane
                if (in_plane) {
                  const int so = c * op.v[20] + (w >> 1) + 1;
                  uint8_t* pe = base + op.v[9];
                  uint8_t* po = base + op.v[10];
                  uint8_t* plane = (w & 1) ? po : pe;
#pragma unroll
                  for (int i4 = 0; i4 < 4; ++i4) {
                    uint4 pk = nk_pack(x + 8 * i4);
                    if (!valid) pk = make_uint4(0u, 0u, 0u, 0u);
                    *nk_chunk(plane, so, hb + i4) = pk;
                    if (!valid) *nk_chunk(po, so, hb + i4) = pk;
                  }
                  if (N == 16) {
#pragma unroll
                    for (int i4 = 4; i4 < 8; ++i4) {
                      *nk_chunk(plane, so, i4) = make_uint4(0u, 0u, 0u, 0u);
                      if (!valid) *nk_chunk(po, so, i4) = make_uint4(0u, 0u, 0u, 0u);
                    }
                  }
                }
              } else if (pj >= 0 && in_plane) {
                uint8_t* plane = base + pj;
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {
                  uint4 pk = nk_pack(x + 8 * i4);
                  if (!valid) pk = make_uint4(0u, 0u, 0u, 0u);          // halo cell of the flat layout
                  *nk_chunk(plane, s, hb + i4) = pk;
                }
                if (N == 16) {               // channels 32..63 of the plane are zero (the next conv reads K = 64)
#pragma unroll
                  for (int i4 = 4; i4 < 8; ++i4) *nk_chunk(plane, s, i4) = make_uint4(0u, 0u, 0u, 0u);
                }
              }
            }
          }
          if (te == 64) {                    // q == 0, lane 0, half 0 (row r = 0 of tile 0): the guard rows of the output planes
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              if (op.v[9 + j] >= 0) {
#pragma unroll
                for (int i4 = 0; i4 < 8; ++i4) *nk_chunk(base + op.v[9 + j], 0, i4) = make_uint4(0u, 0u, 0u, 0u);
              }
            }
          }
          tc_fence_before();
        } else if (type == NK_POOLS) {
          uint8_t* in = base + op.v[1];
          uint8_t* o1 = base + op.v[2];
          uint8_t* o2 = base + op.v[3];
          uint8_t* o3 = base + op.v[4];
          const int R = op.v[5], Wp = op.v[6], W = op.v[7];
          // packed bf16 maxima (HMNMX2.BF16 on the four words of a chunk: the maximum of bf16 values is exact in bf16, nothing to
          // unpack or round): 27 instead of ~160 instructions per tap - the op was issue-bound
          // thread = (chunk j = te & 7, rows s = te / 8 + 32 i): the chunk and the swizzle phase s & 7 are loop invariants and the
          // (clip, column) of the row is carried along instead of divided out (the ops are latency / issue bound)
          const int j = te & 7;
          int r = (te >> 3) - 1, c = 0, w = r;
          for (int s = te >> 3; s <= R; s += NK_EPI / 8, r += NK_EPI / 8, w += NK_EPI / 8) {
            while (w >= Wp) { w -= Wp; ++c; }
            uint4 q1 = make_uint4(0u, 0u, 0u, 0u), q2 = q1, q3 = q1;
            if (r >= 0 && w < W) {
              q1 = q2 = q3 = make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u);      // -inf
#pragma unroll
              for (int d = -6; d <= 6; ++d) {
                const int x = w + d;
                if (x < 0 || x >= W) continue;
                const uint4 v = *nk_chunk(in, c * Wp + x + 1, j);
                const int ad = d < 0 ? -d : d;
                if (ad <= 2) q1 = nk_max_bf16x8(q1, v);
                if (ad <= 4) q2 = nk_max_bf16x8(q2, v);
                q3 = nk_max_bf16x8(q3, v);
              }
            }
            *nk_chunk(o1, s, j) = q1;
            *nk_chunk(o2, s, j) = q2;
            *nk_chunk(o3, s, j) = q3;
          }
        } else if (type == NK_PAIRAVG || type == NK_UP2 || type == NK_DEINT) {
          uint8_t* in = base + op.v[1];
          uint8_t* o1 = base + op.v[2];
          uint8_t* o2 = type == NK_DEINT ? base + op.v[3] : nullptr;
          const int R = op.v[5], Wp = op.v[6], W = op.v[7], Wpi = op.v[8];
          const int j = te & 7;
          int r = (te >> 3) - 1, c = 0, w = r;
          for (int s = te >> 3; s <= R; s += NK_EPI / 8, r += NK_EPI / 8, w += NK_EPI / 8) {
            while (w >= Wp) { w -= Wp; ++c; }
            uint4 ra = make_uint4(0u, 0u, 0u, 0u), rb = ra;
            if (r >= 0 && w < W) {
              if (type == NK_DEINT) {
                ra = *nk_chunk(in, c * Wpi + 2 * w + 1, j);
                rb = *nk_chunk(in, c * Wpi + 2 * w + 2, j);
              } else if (type == NK_PAIRAVG) {
                float a[8], nb[8], v[8];
                nk_unpack(*nk_chunk(in, c * Wpi + 2 * w + 1, j), a);
                nk_unpack(*nk_chunk(in, c * Wpi + 2 * w + 2, j), nb);
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = 0.5f * a[e] + 0.5f * nb[e];
                ra = nk_pack(v);
              } else {
                // bilinear x2 (align_corners = False): 0.75 x nearer + 0.25 x farther source column, one bf16 rounding - as packed
                // bf16 FMAs (0.25 x is exact, the FMA rounds the exact sum once): 12 instead of ~60 instructions per chunk
                const int Wi = W >> 1, k = w >> 1;
                const int k2 = (w & 1) ? (k + 1 < Wi ? k + 1 : Wi - 1) : (k > 0 ? k - 1 : 0);
                ra = nk_lerp_bf16x8(*nk_chunk(in, c * Wpi + k + 1, j), *nk_chunk(in, c * Wpi + k2 + 1, j));
              }
            }
            *nk_chunk(o1, s, j) = ra;
            if (o2 != nullptr) *nk_chunk(o2, s, j) = rb;
          }
        } else if (type == NK_DUMP) {
          if (dbg != nullptr) {
            const uint8_t* plane = base + op.v[1];
            const int rows = op.v[2];
            __nv_bfloat16* dst = dbg + op.v[3] + (int64_t)unit * op.v[4];
            for (int item = te; item < rows * 8; item += NK_EPI) {
              const int s = item >> 3, j = item & 7;
              *reinterpret_cast<uint4*>(dst + (int64_t)s * 64 + 8 * j) = *nk_chunk(const_cast<uint8_t*>(plane), s, j);
            }
          }
        }
        if (tle && te == 0) p.tlog[oi * NK_TL + 3] = clock64();
        if (oi == p.n_ops - 1 || s_ops[oi + 1].v[0] == NK_CONV) {
          fence_proxy_async();   // this thread's plane writes (this op and the element-wise ops before it) become visible to the
          __syncwarp();          // tensor core (async proxy); one arrival per warp (256 arrivals on one mbarrier cost ~500 cycles)
          if (lane == 0) nk_mbar_arrive(op_done);
          if (tle && lane == 0) {
            if (te == 0) p.tlog[oi * NK_TL + 4] = clock64();                                                  // warp 2 arrived
            atomicMax(reinterpret_cast<unsigned long long*>(p.tlog + oi * NK_TL + 5), (unsigned long long)clock64());   // last warp arrived
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, NK_TMEM_COLS);
}

int encode_map_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box);   // conv_tc.cu

int init_neck_fused_attrs() {
  cudaError_t e = cudaFuncSetAttribute(neck_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(neck_fused_kernel) failed: %s", cudaGetErrorString(e));
    return YAD_ERR_CUDA;
  }
  return YAD_OK;
}

}  // namespace yad

static long long* g_nk_tlog = nullptr;
static int g_nk_tl_iter = 0;
/* debug: device buffer of 4 * n_ops int64 that the next launches fill with CTA 0's per-op clock64 stamps of its first clip
 * (MMA warp: start of the op after its dependencies, last MMA issued; epilogue: accumulator complete, op done); NULL switches it off */
extern "C" int yad_neck_fused_set_timeline(void* dev_buf) {
  g_nk_tlog = reinterpret_cast<long long*>(dev_buf);
  return YAD_OK;
}
/* which of CTA 0's clips the timeline records: 0 = its first (cold: nothing prefetched), 1 = its second (steady state), ... */
extern "C" int yad_neck_fused_set_timeline_iter(int32_t iter) {
  g_nk_tl_iter = iter;
  return YAD_OK;
}

extern "C" int yad_neck_fused(const void* const* fmaps, const int32_t* fmap_k, const int32_t* fmap_rows_per_clip,
                              const int32_t* fmap_box_rows, int64_t B, int32_t clips_per_unit,
                              const void* wblob, int64_t wrows, const float* bias, int32_t n_bias, const void* ops, int32_t n_ops,
                              const void* kbs, int32_t n_kb, int32_t pool_bytes, int32_t n_slots, float* const* heads,
                              const int32_t* head_W, int32_t head_ld, void* dbg, yad_stream_t stream) {
  using namespace yad;
  YAD_CHECK_ARG(fmaps && fmap_k && fmap_rows_per_clip && fmap_box_rows && wblob && bias && ops && kbs && heads && head_W,
                "yad_neck_fused: null pointer");
  YAD_CHECK_ARG(clips_per_unit >= 1 && clips_per_unit <= 4, "yad_neck_fused: 1..4 clips per unit");
  YAD_CHECK_ARG(B >= 0 && B < (1 << 22) && wrows >= 16 && n_ops >= 1 && n_ops <= 64 && n_kb >= 1 && n_kb <= 256 && n_bias >= 1,
                "yad_neck_fused: bad sizes");
  YAD_CHECK_ARG(n_slots >= 3 && n_slots <= NK_MAX_SLOTS && pool_bytes >= 1024 && pool_bytes % 1024 == 0, "yad_neck_fused: bad ring / pool");
  YAD_CHECK_ARG(head_ld % 4 == 0 && head_ld >= 4 && head_ld <= 16, "yad_neck_fused: head_ld must be 4..16 floats, a multiple of 4");
  if (B == 0) return YAD_OK;
  const size_t smem = 1024 + (size_t)pool_bytes + (size_t)n_slots * NK_SLOT + (size_t)n_ops * sizeof(NkOp) + (size_t)n_kb * sizeof(NkKb) +
                      8 + (2 * NK_MAX_SLOTS + 2) * 8 + 16;
  YAD_CHECK_ARG(smem <= 227 * 1024, "yad_neck_fused: %zu bytes of shared memory needed", smem);
  CUtensorMap mi[4], mw[3];
  uint32_t a_rows[4];
  for (int i = 0; i < 4; ++i) {
    YAD_CHECK_ARG(fmaps[i] && fmap_k[i] % 64 == 0 && fmap_k[i] >= 64 && fmap_rows_per_clip[i] >= 1 && fmap_box_rows[i] >= 1 &&
                      fmap_box_rows[i] <= 128 && fmap_box_rows[i] % 8 == 0 &&
                      reinterpret_cast<uintptr_t>(fmaps[i]) % 16 == 0,
                  "yad_neck_fused: bad feature map %d", i);
    const uint64_t dims[2] = {(uint64_t)fmap_k[i], (uint64_t)B * (uint64_t)fmap_rows_per_clip[i]};
    const uint64_t strides[1] = {(uint64_t)fmap_k[i] * 2};
    // rows of one A box (the program's choice: a clip's rows rounded up to 8, the clip pitch of a multi-clip unit, or 128); rows
    // of a ring slot past the box keep stale shared-memory contents, which only reach accumulator rows that the epilogue zeroes
    a_rows[i] = (uint32_t)fmap_box_rows[i];
    const uint32_t box[2] = {64u, a_rows[i]};
    int rc = encode_map_bf16(&mi[i], fmaps[i], 2, dims, strides, box);
    if (rc) return rc;
  }
  const uint32_t wn[3] = {16u, 64u, 128u};
  for (int i = 0; i < 3; ++i) {
    const uint64_t dims[2] = {64u, (uint64_t)wrows};
    const uint64_t strides[1] = {128u};
    const uint32_t box[2] = {64u, wn[i]};
    int rc = encode_map_bf16(&mw[i], wblob, 2, dims, strides, box);
    if (rc) return rc;
  }
  NkParams p;
  p.tlog = g_nk_tlog;
  p.n_ops = n_ops;
  p.n_kb = n_kb;
  p.n_clips = (int)B;
  p.G = clips_per_unit;
  p.n_units = (int)((B + clips_per_unit - 1) / clips_per_unit);
  p.n_slots = n_slots;
  p.pool_bytes = pool_bytes;
  p.n_bias = n_bias;
  for (int i = 0; i < 3; ++i) p.head_W[i] = head_W[i];
  p.head_ld = head_ld;
  for (int i = 0; i < 4; ++i) p.a_bytes[i] = (int32_t)(a_rows[i] * 128u);
  p.tl_iter = g_nk_tl_iter;
  const int nsm = sm_count() > 0 ? sm_count() : 148;
  const unsigned grid = (unsigned)(p.n_units < nsm ? p.n_units : nsm);
  YAD_CUDA(launch_pdl(neck_fused_kernel, dim3(grid), dim3(NK_THREADS), smem, (cudaStream_t)stream, mi[0], mi[1], mi[2], mi[3], mw[0], mw[1],
                      mw[2], p, reinterpret_cast<const NkOp*>(ops), reinterpret_cast<const NkKb*>(kbs), bias, heads[0], heads[1], heads[2],
                      reinterpret_cast<__nv_bfloat16*>(dbg)));
  return YAD_OK;
}
