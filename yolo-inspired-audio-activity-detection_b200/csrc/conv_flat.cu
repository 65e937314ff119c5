// Patch-resident tcgen05 / TMEM implicit-GEMM convolution for stride-1 convs on "flat" (halo-padded) activations.
//
// Replaces F.conv2d + folded BatchNorm + activation (+ residual) for the 3x3 stride-1 convolutions of the
// reference's ResNet stages (torchvision resnet.py:89-105 via modules/_backbone.py:148-151).
//
// Why a second conv kernel: the tap-by-tap implicit GEMM of conv_tc.cu re-reads every input element from L2 once per
// filter tap (9x for a 3x3); with Cout = 64..128 that makes the layer L2-bandwidth bound at ~1/4 of the tensor peak.
// Here the activations live in HBM in a flat, halo-padded layout
//     f = (b * Wp + w) * Hp + h        (h fastest; cells with h >= H or w >= W are zero and are never written)
// so that for a stride-1 conv EVERY filter tap is a constant shift in f:  in[f + dw * Hp + dh].  One CTA therefore
// loads a contiguous patch of pixels [f0 + min_off, f0 + 128 * MT + max_off) ONCE per 64-channel chunk (plain 2-D
// TMA, 128-byte swizzled rows) and issues the MMAs of all taps from it by shifting the start address of the A-operand
// shared-memory descriptor by whole 128-byte rows.  Measured on B200: the tensor core applies the 128B-swizzle XOR to
// the absolute shared-memory address bits, exactly like the TMA unit that wrote the patch, so a start address on any
// 128-byte row reads the right data with the descriptor's base_offset field left at 0 (setting it to (addr >> 7) & 7
// scrambled the rows: 7 of 8 parity cases failed with it and all pass without, round-1 logs), at no cost in MMA rate
// (tools/micro/mma_rate.cu: 48 / 64 / 128 cycles per MMA at N = 64 / 128 / 256 with or without the row shift).
// Because h is the fastest index and H is small (8, 4, 2, 1), the halo of a 3x3 is only Hp + 1 pixels.
// Weights stream through their own ring, one [BN x 64] block per (tap, chunk), shared by the MT accumulators.
//
// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread tcgen05.mma issuer,
// warps 2..9 = epilogue (tcgen05.ld -> bias / residual / activation; two warps per TMEM lane quadrant, each taking every
// second 32-column chunk), warp 10 = epilogue TMA (residual tiles in, output tiles out), TMEM accumulators double-buffered
// so the epilogue of super-tile i overlaps the MMAs of super-tile i + 1.
//
// Epilogue memory traffic goes through the TMA unit: a thread owns one pixel row of the accumulator, so register-sourced
// stores / residual loads are 16 bytes per lane at a 128- or 256-byte lane stride - one LSU wavefront per 16 bytes.  ncu on
// the layer1 shape (profiles/r02_ncu_conv_flat_l1_*): l1tex__data_pipe_lsu_wavefronts at 71 % of peak, 135 k global
// wavefronts per SM, tensor pipe 34 %: the kernel was bound by its own epilogue.  Now each [128 pixel x 64 channel] tile of
// the residual is bulk-loaded into a swizzled 16 KB staging buffer, the epilogue warps add / activate / round IN PLACE
// (conflict-free 16-byte shared-memory accesses) and the tile is bulk-stored from the same buffer; halo pixels are written
// as the zeros they have to stay.  Three staging buffers: the residual tile of item g + 2 lands while item g is computed.
#include "common.cuh"
#include "tc_ptx.cuh"
#include <stdlib.h>
#include <string.h>

namespace yad {

// Warps: 0 weight TMA, 1 MMA stream A, 2..9 epilogue, 10 epilogue TMA (stores / residual loads), 11 patch TMA, 12 MMA stream B
constexpr int FL_THREADS = 416;
constexpr int FL_SRC2 = 0x4000;
constexpr int FL_NSTAGE = 3;        // epilogue staging buffers of [128 rows x 128 B]
constexpr int FL_STAGE_BYTES = 128 * 128;
constexpr int FL_EPI_THREADS = 256;
constexpr int FL_EPI_WARPS = FL_EPI_THREADS / 32;
constexpr int FL_MAX_STEPS = 64;
constexpr int FL_MAX_RING = 8;      // weight ring slots
constexpr int FL_MAX_NA = 4;        // patch slots per stream
constexpr int FL_BOX_ROWS = 32;     // pixel rows per TMA box of the patch (4 KB)
constexpr int FL_ACC_COLS = 128;    // TMEM columns of one accumulator stage of one stream (MT x BN)

struct FlatParams {
  int64_t F;                    // flat pixels = B * Wp * Hp
  int32_t H, W, Hp, Wp;
  int32_t Hpo, Wpo;             // output / residual pitches (same image size)
  int32_t remap;                // 1 when (Hpo, Wpo) != (Hp, Wp)
  int32_t MT, BN, n_ntiles, n_super;   // tile = 128 * MT pixel rows x BN channels (MT * BN = 128); n_super = M tiles * n_ntiles
  int32_t n_steps;
  int32_t NA, NW;               // ring depths: patch slots per stream, weight slots
  int32_t patch_rows, patch_bytes, w_bytes;
  int32_t min_off;
  int32_t Cout, ld_out, co_off, ld_res, act;
  uint32_t idesc;
  int32_t flags;                // bit 4: fine timeline of the MMA warps (with tlog)
  int32_t tma_epi;              // 1: residual / output tiles move through the TMA unit and the staging buffers (needs Cout % 64 == 0,
                                // same pitches in and out); 0: register-sourced global accesses
  uint32_t hp_mul, hp_sh, wp_mul, wp_sh;   // f / Hp and col / Wp as multiply-high (exact for f < 2^31): (f * mul) >> sh
  int32_t s2d_Hp, s2d_Wp;       // second, space-to-depth copy of the output (0: none): pixel (b, h, w) goes to flat cell
                                // (b * s2d_Wp + w / 2) * s2d_Hp + h / 2, channel plane (h & 1) * 2 + (w & 1) of 4 * Cout channels
  uint32_t step_mma[FL_MAX_STEPS];    // MMA warp: bits 0..15 = first patch row of the tap in 16-byte units, bit 30 = first
                                      // step of its chunk, bit 31 = last step of its chunk
  int16_t step_off[FL_MAX_STEPS];     // flat pixel shift of the tap
  int16_t step_chunk[FL_MAX_STEPS];   // 64-channel chunk; bit 14 (FL_SRC2): of the SECOND input tensor (same flat geometry)
  int8_t step_first[FL_MAX_STEPS];    // first step of its chunk (load a new patch)
  int8_t step_last[FL_MAX_STEPS];     // last step of its chunk (release the patch)
  int32_t step_wk[FL_MAX_STEPS];      // K offset of the [BN x 64] weight block
  long long* tlog;                    // debug timeline (16 int64 per CTA), NULL in production
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// exact n / d for n < 2^31 with host-computed (mul, sh) = (ceil(2^(31 + s) / d), 31 + s), 2^s >= d
__device__ __forceinline__ uint32_t fast_div(uint32_t n, uint32_t mul, uint32_t sh) { return (uint32_t)(((uint64_t)n * mul) >> sh); }
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// K-major SWIZZLE_128B shared-memory descriptor = (FL_DESC_HI << 32) | lo, lo = (address >> 4) | LBO field (1 << 16);
// high word: SBO = 1024 B (8 rows x 128 B), descriptor version 1, SWIZZLE_128B, base_offset 0.  The start address may sit
// on any 128-byte row of a swizzle atom (see the header comment): shifting by r rows adds 8 * r to lo, by 16 K-elements 2.
constexpr uint32_t FL_DESC_HI = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ uint64_t desc64(uint32_t lo) { return ((uint64_t)FL_DESC_HI << 32) | (uint64_t)lo; }

// One MMA-issuing warp = one STREAM.  Measured on B200 (tools/micro/mma_queue.cu, profiles/r02_mma_queue.txt): the tensor pipe
// queues ONE tcgen05.mma behind the executing one (less after a tcgen05.commit), so whatever the issuing thread does between
// two steps - step-table load, mbarrier polls, fence, elect, a dozen R2UR, commits: ~210 cycles - runs with an idle pipe:
// a lone issuing warp reached 91 / 69 cycles per MMA at N = 128 / 64 against 64 / 48 back to back, with no memory traffic and
// no epilogue at all (YAD_FLAT_DBG experiments of round 2).  Two warps issuing into DISJOINT accumulators hide each other's
// gaps completely (same micro-benchmark: 2 x (4 MMAs + up to ~300 idle cycles) per 512 cycles).  So the CTA runs two streams:
// stream A takes the even, stream B the odd tiles of the CTA's tile sequence, each with its own patch ring and its own
// double-buffered accumulator (TMEM columns (2 * stream + stage) * 128); both consume the SAME weight blocks (the tiles of a
// CTA share one N tile), whose ring slots are released by one commit of each stream.
// The whole warp walks the (warp-uniform) loops so that descriptors stay in uniform registers; one elected lane issues.
template <int MT, int BN, bool TL, bool PAIR>
__device__ __forceinline__ void flat_mma_loop(const FlatParams& p, const int stream, const int n_my, uint8_t* sm_a, uint8_t* sm_w,
                                              uint64_t* full_a, uint64_t* empty_a, uint64_t* full_w, uint64_t* empty_w,
                                              uint64_t* tmem_full, uint64_t* tmem_empty, long long* tl) {
  uint32_t a_slot = 0, a_phase = 0, w_slot = 0, w_phase = 0, as = 0, acc_phase = 0, a_cur = 0, a_lo0 = 0;
  const uint32_t w_lo0 = desc_lo(smem_u32(sm_w));
  const uint32_t a_base_lo = desc_lo(smem_u32(sm_a));
  const uint32_t a_slot_units = (uint32_t)p.patch_bytes >> 4;
  constexpr uint32_t W_UNITS = (PAIR ? BN / 2 : BN) * 128 / 16;      // a CTA of a pair holds half of the weight block's rows
  const int n_steps = p.n_steps;
  const uint32_t idesc = p.idesc;
  long long t_begin = 0, t_acc = 0, t_a = 0, t_w = 0, t_issue = 0, t0 = 0, t1 = 0;
  if (TL) t_begin = clock64();
  for (int j = 0; j < n_my; ++j) {                  // one tile of this stream per tile pair
    if (TL) t0 = clock64();
    mbar_wait(&tmem_empty[as], acc_phase ^ 1);      // epilogue drained this accumulator stage (free on the first lap)
    tc_fence_after();
    if (TL) { t1 = clock64(); t_acc += t1 - t0; }
    const uint32_t d0 = (uint32_t)(2 * stream + as) * FL_ACC_COLS;
    for (int s = 0; s < n_steps; ++s) {
      const uint32_t sm = p.step_mma[s];
      if (sm & (1u << 30)) {
        a_cur = a_slot;
        if (TL) t0 = clock64();
        mbar_wait(&full_a[a_cur], a_phase);
        if (TL) { t1 = clock64(); t_a += t1 - t0; }
        a_lo0 = a_base_lo + a_cur * a_slot_units;
        if (++a_slot == (uint32_t)p.NA) { a_slot = 0; a_phase ^= 1; }
      }
      if (TL) t0 = clock64();
      mbar_wait(&full_w[w_slot], w_phase);
      tc_fence_after();
      if (TL) { t1 = clock64(); t_w += t1 - t0; }
      if (elect_one()) {
        const uint32_t a_lo = a_lo0 + (sm & 0xFFFFu);
        const uint32_t b_lo = w_lo0 + w_slot * W_UNITS;
        const uint32_t acc0 = s > 0 ? 1u : 0u;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (PAIR)
              umma_bf16_pair(d0 + mt * BN, desc64(a_lo + mt * 1024 + 2 * k), desc64(b_lo + 2 * k), idesc, k > 0 ? 1u : acc0);
            else
              umma_bf16(d0 + mt * BN, desc64(a_lo + mt * 1024 + 2 * k), desc64(b_lo + 2 * k), idesc, k > 0 ? 1u : acc0);
          }
        }
        if (PAIR) {                                   // the same barriers of both CTAs
          umma_commit_pair(&empty_w[w_slot]);
          if (sm & (1u << 31)) umma_commit_pair(&empty_a[a_cur]);
          if (s == n_steps - 1) umma_commit_pair(&tmem_full[as]);
        } else {
          umma_commit(&empty_w[w_slot]);
          if (sm & (1u << 31)) umma_commit(&empty_a[a_cur]);
          if (s == n_steps - 1) umma_commit(&tmem_full[as]);
        }
      }
      __syncwarp();
      if (TL) t_issue += clock64() - t1;
      if (++w_slot == (uint32_t)p.NW) { w_slot = 0; w_phase ^= 1; }
    }
    if ((as ^= 1) == 0) acc_phase ^= 1;
  }
  if (TL && stream == 0 && (threadIdx.x & 31) == 0) {
    tl[0] = clock64() - t_begin;
    tl[1] = t_acc;
    tl[2] = t_a;
    tl[3] = t_w;
    tl[4] = t_issue;
  }
}

template <bool PAIR>
__global__ void __launch_bounds__(FL_THREADS, 1)
conv_flat_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a2,
                 const __grid_constant__ CUtensorMap map_w,
                 const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_res,
                 const __grid_constant__ FlatParams p, const float* __restrict__ bias,
                 const __nv_bfloat16* __restrict__ residual, __nv_bfloat16* __restrict__ out,
                 __nv_bfloat16* __restrict__ out_s2d) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sm_a = smem;                                       // [2 streams][NA][patch_bytes]
  uint8_t* sm_w = sm_a + (size_t)2 * p.NA * p.patch_bytes;    // [NW][w_bytes]
  uint8_t* sm_o = sm_w + (size_t)p.NW * p.w_bytes;            // [FL_NSTAGE][16 KB] epilogue staging (1024-byte aligned: patch and
                                                              // weight slots are multiples of 1 KB)
  uint64_t* full_a = reinterpret_cast<uint64_t*>(sm_o + (p.tma_epi ? FL_NSTAGE * FL_STAGE_BYTES : 0));   // [2][FL_MAX_NA]
  uint64_t* empty_a = full_a + 2 * FL_MAX_NA;                 // [2][FL_MAX_NA]
  uint64_t* full_w = empty_a + 2 * FL_MAX_NA;                 // [FL_MAX_RING]
  uint64_t* empty_w = full_w + FL_MAX_RING;                   // [FL_MAX_RING]
  uint64_t* tmem_full = empty_w + FL_MAX_RING;                // [2 streams][2 stages]
  uint64_t* tmem_empty = tmem_full + 4;                       // [2 streams][2 stages]
  uint64_t* buf_ready = tmem_empty + 4;                       // [FL_NSTAGE] staging buffer free (and its residual tile landed)
  uint64_t* out_full = buf_ready + FL_NSTAGE;                 // [FL_NSTAGE] staging buffer holds a finished output tile
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(out_full + FL_NSTAGE);
  float* s_bias = reinterpret_cast<float*>(tmem_ptr_smem + 4);   // [n_ntiles * BN]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ncout_pad = p.n_ntiles * p.BN;
  // coarse debug timeline (p.tlog, 16 int64 per CTA): [8] entry, [9] set-up done, [10] / [11] MMA loop of stream A, [12] epilogue
  // done, [13] exit (clock64), [14] / [15] globaltimer at entry / exit; [0..4] see flat_mma_loop
  long long* tl = p.tlog != nullptr ? p.tlog + (size_t)blockIdx.x * 16 : nullptr;
  if (tl != nullptr && threadIdx.x == 32) {
    tl[8] = clock64();
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    tl[14] = (long long)gt;
  }
  for (int i = threadIdx.x; i < ncout_pad; i += FL_THREADS) s_bias[i] = bias[i];

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_w);
    for (int s = 0; s < 2 * FL_MAX_NA; ++s) {
      mbar_init(&full_a[s], 1);
      mbar_init(&empty_a[s], 1);
    }
    for (int s = 0; s < FL_MAX_RING; ++s) {
      mbar_init(&full_w[s], 1);
      mbar_init(&empty_w[s], 2);             // one commit of each stream
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], PAIR ? 2 * FL_EPI_WARPS : FL_EPI_WARPS);     // one arrival per epilogue warp (of both CTAs)
    }
    for (int s = 0; s < FL_NSTAGE; ++s) {
      mbar_init(&buf_ready[s], 1);
      mbar_init(&out_full[s], FL_EPI_WARPS);     // one arrival per warp: 256 arrivals on one mbarrier serialise (~500 cycles)
    }
    if (p.tma_epi) {
      prefetch_tmap(&map_out);
      if (residual != nullptr) prefetch_tmap(&map_res);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR)
      tmem_alloc_pair(tmem_ptr_smem, 512);
    else
      tmem_alloc(tmem_ptr_smem, 512);
  }
  tc_fence_before();
  if (PAIR)
    cluster_sync_all();        // the peer's barriers are initialised before any TMA completion / commit / arrive reaches them
  else
    __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();          // programmatic dependent launch: the prologue above overlapped the previous kernel's tail
  pdl_trigger();
  if (tl != nullptr && threadIdx.x == 32) tl[9] = clock64();
  const int rows_per_super = 128 * p.MT;
  const int n_boxes = p.patch_rows / FL_BOX_ROWS;
  // Work units: a CTA (a CTA pair with PAIR) walks the sequence of tile PAIRS P = unit + j * n_units, j < n_my; pair P = N tile
  // P % n_ntiles, M-pair P / n_ntiles; its two tiles (stream 0 / 1) are adjacent M tiles and share every weight block.  With PAIR a
  // tile is 2 x 128 * MT rows, one half per CTA (cluster rank).  Tiles past the end of the tensor are computed on zero-filled
  // patches and dropped by the epilogue's masks / the TMA unit's clipping.
  const int rank = PAIR ? (int)cluster_ctarank() : 0;
  const int unit = PAIR ? (int)blockIdx.x >> 1 : (int)blockIdx.x;
  const int n_units = PAIR ? (int)gridDim.x >> 1 : (int)gridDim.x;
  const int n_my = unit < p.n_super ? (p.n_super - unit + n_units - 1) / n_units : 0;
  auto pair_nt = [&](int j) { return (unit + j * n_units) % p.n_ntiles; };
  auto tile_m = [&](int j, int q) {            // M tile (of 128 * MT rows) of stream q of this CTA in pair j
    const int mp = (unit + j * n_units) / p.n_ntiles;
    return PAIR ? (2 * mp + q) * 2 + rank : 2 * mp + q;
  };
  const uint32_t leader_full_a = PAIR ? mapa_u32(full_a, 0) : 0u, leader_full_w = PAIR ? mapa_u32(full_w, 0) : 0u;
  const uint32_t leader_tmem_empty = PAIR ? mapa_u32(tmem_empty, 0) : 0u;

  if (warp == 0) {
    // ===================== weight producer: one block per step of every tile PAIR =====================
    // Patches and weights have their own producer threads (this warp and warp 11): a patch is requested the moment its slot is
    // released, not after the NW weight blocks in front of it.
    if (lane == 0) {
      // ring state kept as (slot, phase) counters: no integer division on this latency-bound single-thread path.
      // Waiting on parity (phase ^ 1) of a fresh mbarrier returns at once, so the first lap needs no special case.
      uint32_t w_slot = 0, w_phase = 0;
      const uint32_t wb = (uint32_t)p.w_bytes;          // this CTA's share of a weight block (half of its rows with PAIR)
      for (int j = 0; j < n_my; ++j) {
        const int n0 = pair_nt(j) * p.BN + (PAIR ? rank * (p.BN >> 1) : 0);
        for (int s = 0; s < p.n_steps; ++s) {
          mbar_wait(&empty_w[w_slot], w_phase ^ 1);
          if (PAIR) {
            // both CTAs' halves complete on the LEADER's barrier (its MMA warps wait there); only the leader announces the bytes
            if (rank == 0) mbar_expect_tx(&full_w[w_slot], 2u * wb);
            tma_load_2d_pair(&map_w, leader_full_w + 8u * w_slot, sm_w + (size_t)w_slot * wb, p.step_wk[s], n0);
          } else {
            mbar_expect_tx(&full_w[w_slot], wb);
            tma_load_2d(&map_w, &full_w[w_slot], sm_w + (size_t)w_slot * wb, p.step_wk[s], n0);
          }
          if (++w_slot == (uint32_t)p.NW) { w_slot = 0; w_phase ^= 1; }
        }
      }
    }
  } else if (warp == 11) {
    // ===================== patch producer: per chunk of a tile pair, stream A's patch then stream B's =====================
    // (the streams advance in lock step - they share the weight ring - so serving them in program order cannot starve one)
    if (lane == 0) {
      uint32_t a_slot[2] = {0, 0}, a_phase[2] = {0, 0};
      for (int j = 0; j < n_my; ++j) {
        for (int s = 0; s < p.n_steps; ++s) {
          if (!p.step_first[s]) continue;
          const int c0 = (p.step_chunk[s] & (FL_SRC2 - 1)) * 64;
          const CUtensorMap* ma = (p.step_chunk[s] & FL_SRC2) ? &map_a2 : &map_a;
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int r0 = (int)((int64_t)tile_m(j, q) * rows_per_super + p.min_off);
            const int bi = q * FL_MAX_NA + (int)a_slot[q];
            mbar_wait(&empty_a[bi], a_phase[q] ^ 1);
            uint8_t* dst = sm_a + (size_t)(q * p.NA + a_slot[q]) * p.patch_bytes;
            if (PAIR) {
              if (rank == 0) mbar_expect_tx(&full_a[bi], 2u * (uint32_t)p.patch_bytes);
              for (int b = 0; b < n_boxes; ++b)
                tma_load_2d_pair(ma, leader_full_a + 8u * bi, dst + b * (FL_BOX_ROWS * 128), c0, r0 + b * FL_BOX_ROWS);
            } else {
              mbar_expect_tx(&full_a[bi], (uint32_t)p.patch_bytes);
              for (int b = 0; b < n_boxes; ++b) tma_load_2d(ma, &full_a[bi], dst + b * (FL_BOX_ROWS * 128), c0, r0 + b * FL_BOX_ROWS);
            }
            if (++a_slot[q] == (uint32_t)p.NA) { a_slot[q] = 0; a_phase[q] ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1 || warp == 12) {
    // ===================== MMA issuers: stream A (warp 1), stream B (warp 12) =====================
    // All 512 TMEM columns are ours, so the allocation starts at column 0 and accumulator addresses are plain constants.
    if (tmem_base != 0) __trap();
    const int stream = warp == 1 ? 0 : 1;
    uint8_t* sa = sm_a + (size_t)stream * p.NA * p.patch_bytes;
    uint64_t* fa = full_a + stream * FL_MAX_NA;
    uint64_t* ea = empty_a + stream * FL_MAX_NA;
    uint64_t* tf = tmem_full + 2 * stream;
    uint64_t* te = tmem_empty + 2 * stream;
    if (tl != nullptr && warp == 1 && lane == 0) tl[10] = clock64();
    if (rank == 0) {           // with PAIR only the leader CTA issues (for both CTAs)
      if (p.tlog != nullptr && (p.flags & 16)) {
        if (p.MT == 2)
          flat_mma_loop<2, 64, true, PAIR>(p, stream, n_my, sa, sm_w, fa, ea, full_w, empty_w, tf, te, tl);
        else
          flat_mma_loop<1, 128, true, PAIR>(p, stream, n_my, sa, sm_w, fa, ea, full_w, empty_w, tf, te, tl);
      } else if (p.MT == 2) {
        flat_mma_loop<2, 64, false, PAIR>(p, stream, n_my, sa, sm_w, fa, ea, full_w, empty_w, tf, te, tl);
      } else {
        flat_mma_loop<1, 128, false, PAIR>(p, stream, n_my, sa, sm_w, fa, ea, full_w, empty_w, tf, te, tl);
      }
    }
    if (tl != nullptr && warp == 1 && lane == 0) tl[11] = clock64();
  } else if (warp == 10) {
    // ===================== epilogue TMA: stores the finished tiles, loads the residual tiles two items ahead =====================
    if (p.tma_epi && lane == 0) {
      const int cpa = p.BN >> 6, n_items = p.MT * cpa;
      const int total = 2 * n_my * n_items;
      // item g of this CTA -> (channel, row) coordinate of its [128 x 64] tile; tiles in sequence order i = 2 j + stream
      auto coords = [&](int g, int& ch, int& row) {
        const int i = g / n_items, jj = g - i * n_items;
        const int mt = jj / cpa, cj = jj - mt * cpa;
        ch = pair_nt(i >> 1) * p.BN + 64 * cj;
        row = tile_m(i >> 1, i & 1) * rows_per_super + 128 * mt;
      };
      auto make_ready = [&](int g) {          // buffer g % FL_NSTAGE is free: fetch the residual tile of item g, or just say so
        uint64_t* bar = &buf_ready[g % FL_NSTAGE];
        if (residual != nullptr) {
          int ch, row;
          coords(g, ch, row);
          mbar_expect_tx(bar, FL_STAGE_BYTES);
          tma_load_2d(&map_res, bar, sm_o + (g % FL_NSTAGE) * FL_STAGE_BYTES, ch, row);
        } else {
          mbar_arrive(bar);
        }
      };
      for (int g = 0; g < total && g < FL_NSTAGE - 1; ++g) make_ready(g);
      for (int g = 0; g < total; ++g) {
        const int b = g % FL_NSTAGE;
        mbar_wait(&out_full[b], (uint32_t)(g / FL_NSTAGE) & 1u);
        int ch, row;
        coords(g, ch, row);
        tma_store_2d(&map_out, sm_o + b * FL_STAGE_BYTES, p.co_off + ch, row);
        bulk_commit_group();
        if (g + FL_NSTAGE - 1 < total) {
          bulk_wait_group_read<1>();          // the store of item g - 1 has read its buffer: it is the one item g + 2 uses
          make_ready(g + FL_NSTAGE - 1);
        }
      }
      bulk_wait_group<0>();                   // all tiles written before the CTA (and its shared memory) goes away
    }
  } else {
    // ===================== epilogue: 8 warps; TMEM lane quadrant = warp % 4, column-chunk parity = (warp - 2) / 4 =====
    // Tiles are drained in sequence order i = 0, 1, 2, ...: stream i & 1, accumulator stage (i >> 1) & 1, barrier phase (i >> 2) & 1
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int chunks_per_acc = p.BN >> 6;            // 32-column chunks handled by this warp per accumulator (1 or 2)
    const int n_items = p.MT * chunks_per_acc;       // work items per tile: (mt, chunk)
    const uint32_t Hp = (uint32_t)p.Hp, Wp = (uint32_t)p.Wp;
    const int cpa_sh = chunks_per_acc >> 1;          // chunks_per_acc is 1 or 2
    const uint32_t s_bias_addr = smem_u32(s_bias);
    if (p.tma_epi) {
      // ---- tiles through the staging buffers (see the header): item = (mt, 64-channel block) = one [128 x 64] tile that the 8
      //      warps fill together; thread = (pixel row r, channel half): four 16-byte chunks of its row, swizzled like the TMA image
      uint32_t sb = 0, sb_phase = 0;
      const uint32_t row_off = (uint32_t)r * 128u, rx = (uint32_t)r & 7u;
      const bool etl = tl != nullptr && (p.flags & 16) && threadIdx.x == 64;     // fine timeline of one epilogue thread
      long long e_full = 0, e_buf = 0, e_ld = 0, e0 = 0;
      for (int i = 0; i < 2 * n_my; ++i) {
        const int mtile = tile_m(i >> 1, i & 1);
        const int n0 = pair_nt(i >> 1) * p.BN;
        const uint32_t acc = (uint32_t)(2 * (i & 1) + ((i >> 1) & 1));      // accumulator (stream, stage) = barrier index
        const uint32_t fbase = (uint32_t)mtile * (uint32_t)rows_per_super + (uint32_t)r;
        if (etl) e0 = clock64();
        mbar_wait(&tmem_full[acc], (uint32_t)(i >> 2) & 1u);
        tc_fence_after();
        if (etl) e_full += clock64() - e0;
        for (int j = 0; j < n_items; ++j) {
          const int mt = j >> cpa_sh, cj = j & (chunks_per_acc - 1);
          const int c0 = 32 * (2 * cj + half);
          const uint32_t f = fbase + 128u * (uint32_t)mt;
          const uint32_t col = fast_div(f, p.hp_mul, p.hp_sh), h = f - col * Hp, bb = fast_div(col, p.wp_mul, p.wp_sh), w = col - bb * Wp;
          const bool ok = (int64_t)f < p.F && h < (uint32_t)p.H && w < (uint32_t)p.W;
          const int nbase = n0 + c0;
          uint32_t v[32];
          if (etl) e0 = clock64();
          tmem_ld32(((uint32_t)(q * 32) << 16) + acc * FL_ACC_COLS + (uint32_t)(mt * p.BN + c0), v);
          tmem_ld_wait();
          if (etl) e_ld += clock64() - e0;
          uint8_t* S = sm_o + sb * FL_STAGE_BYTES + row_off;
          if (etl) e0 = clock64();
          mbar_wait(&buf_ready[sb], sb_phase);       // the buffer is free and (if any) the residual tile has landed in it
          if (etl) e_buf += clock64() - e0;
          uint4 pk[4];
          if (ok) {
            float x[32];
            const uint32_t bp = s_bias_addr + (uint32_t)nbase * 4u;
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 b4 = lds_f4(bp + 16u * j4);
              x[j4 * 4 + 0] = __uint_as_float(v[j4 * 4 + 0]) + b4.x;
              x[j4 * 4 + 1] = __uint_as_float(v[j4 * 4 + 1]) + b4.y;
              x[j4 * 4 + 2] = __uint_as_float(v[j4 * 4 + 2]) + b4.z;
              x[j4 * 4 + 3] = __uint_as_float(v[j4 * 4 + 3]) + b4.w;
            }
            if (residual != nullptr) {
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) {
                const uint4 rq = *reinterpret_cast<const uint4*>(S + ((((uint32_t)(4 * half + j4)) ^ rx) << 4));
                const uint32_t ww[4] = {rq.x, rq.y, rq.z, rq.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  x[j4 * 8 + e * 2 + 0] += __uint_as_float(ww[e] << 16);
                  x[j4 * 8 + e * 2 + 1] += __uint_as_float(ww[e] & 0xffff0000u);
                }
              }
            }
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) x[jj] = apply_act(x[jj], p.act);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              uint32_t ww[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                __nv_bfloat162 h2 = __floats2bfloat162_rn(x[j4 * 8 + e * 2], x[j4 * 8 + e * 2 + 1]);
                ww[e] = *reinterpret_cast<uint32_t*>(&h2);
              }
              pk[j4] = make_uint4(ww[0], ww[1], ww[2], ww[3]);
            }
            if (p.s2d_Hp) {
              uint4* op2 = reinterpret_cast<uint4*>(
                  out_s2d + (((bb * (uint32_t)p.s2d_Wp + (w >> 1)) * (uint32_t)p.s2d_Hp + (h >> 1)) * 4u + ((h & 1u) << 1 | (w & 1u))) *
                                (uint32_t)p.Cout + (uint32_t)nbase);
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) op2[j4] = pk[j4];
            }
          } else {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) pk[j4] = make_uint4(0u, 0u, 0u, 0u);      // halo cell: stays zero in global memory
          }
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) *reinterpret_cast<uint4*>(S + ((((uint32_t)(4 * half + j4)) ^ rx) << 4)) = pk[j4];
          fence_proxy_async();               // generic-proxy writes -> visible to the bulk store
          __syncwarp();
          if (lane == 0) mbar_arrive(&out_full[sb]);
          if (++sb == FL_NSTAGE) { sb = 0; sb_phase ^= 1; }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR)
            mbar_arrive_cluster(leader_tmem_empty + 8u * acc);
          else
            mbar_arrive(&tmem_empty[acc]);
        }
      }
      if (etl) {
        tl[5] = e_full;
        tl[6] = e_buf;
        tl[7] = e_ld;
      }
    } else {
      for (int i = 0; i < 2 * n_my; ++i) {
        const int mtile = tile_m(i >> 1, i & 1);
        const int n0 = pair_nt(i >> 1) * p.BN;
        const uint32_t acc = (uint32_t)(2 * (i & 1) + ((i >> 1) & 1));
        const uint32_t fbase = (uint32_t)mtile * (uint32_t)rows_per_super + (uint32_t)r;
        // residual prefetch for item 0 (independent of the accumulator): hides the HBM latency behind the barrier wait
        uint4 rq[4];
        // item j -> (output flat index, first channel, valid); the output may use other pitches than the input
        uint32_t f2_cur = 0;          // element offset of the item's 32 channels in the space-to-depth copy
        auto item_geom = [&](int j, uint32_t& f, int& nbase, bool& ok) {
          const int mt = j / chunks_per_acc, cj = j - mt * chunks_per_acc;
          f = fbase + 128u * (uint32_t)mt;
          const uint32_t col = f / Hp, h = f - col * Hp, bb = col / Wp, w = col - bb * Wp;
          nbase = n0 + 32 * (2 * cj + half);
          ok = (int64_t)f < p.F && h < (uint32_t)p.H && w < (uint32_t)p.W && nbase < p.Cout;
          if (p.remap) f = (bb * (uint32_t)p.Wpo + w) * (uint32_t)p.Hpo + h;
          if (p.s2d_Hp)
            f2_cur = (((bb * (uint32_t)p.s2d_Wp + (w >> 1)) * (uint32_t)p.s2d_Hp + (h >> 1)) * 4u + ((h & 1u) << 1 | (w & 1u))) * (uint32_t)p.Cout +
                     (uint32_t)nbase;
        };
        auto load_res = [&](uint32_t f, int nbase, bool ok) {
          if (residual != nullptr && ok) {
            const uint4* rp = reinterpret_cast<const uint4*>(residual + (int64_t)f * p.ld_res + nbase);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) rq[j4] = __ldg(rp + j4);
          }
        };
        uint32_t f_cur;
        int nb_cur;
        bool ok_cur;
        item_geom(0, f_cur, nb_cur, ok_cur);
        load_res(f_cur, nb_cur, ok_cur);
        mbar_wait(&tmem_full[acc], (uint32_t)(i >> 2) & 1u);
        tc_fence_after();
        for (int j = 0; j < n_items; ++j) {
          const int mt = j / chunks_per_acc, cj = j - mt * chunks_per_acc;
          const int c0 = 32 * (2 * cj + half);
          uint32_t v[32];
          tmem_ld32(((uint32_t)(q * 32) << 16) + acc * FL_ACC_COLS + (uint32_t)(mt * p.BN + c0), v);
          tmem_ld_wait();
          float x[32];
          if (ok_cur) {
            const float4* bp = reinterpret_cast<const float4*>(s_bias + nb_cur);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 b4 = bp[j4];
              x[j4 * 4 + 0] = __uint_as_float(v[j4 * 4 + 0]) + b4.x;
              x[j4 * 4 + 1] = __uint_as_float(v[j4 * 4 + 1]) + b4.y;
              x[j4 * 4 + 2] = __uint_as_float(v[j4 * 4 + 2]) + b4.z;
              x[j4 * 4 + 3] = __uint_as_float(v[j4 * 4 + 3]) + b4.w;
            }
            if (residual != nullptr) {
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) {
                const uint32_t ww[4] = {rq[j4].x, rq[j4].y, rq[j4].z, rq[j4].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  x[j4 * 8 + e * 2 + 0] += __uint_as_float(ww[e] << 16);
                  x[j4 * 8 + e * 2 + 1] += __uint_as_float(ww[e] & 0xffff0000u);
                }
              }
            }
          }
          const uint32_t f_st = f_cur, f2_st = f2_cur;
          const int nb_st = nb_cur;
          const bool ok_st = ok_cur;
          if (j + 1 < n_items) {       // prefetch the next item's residual before the stores of this one
            item_geom(j + 1, f_cur, nb_cur, ok_cur);
            load_res(f_cur, nb_cur, ok_cur);
          }
          if (ok_st) {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) x[jj] = apply_act(x[jj], p.act);
            uint4* op = reinterpret_cast<uint4*>(out + (int64_t)f_st * p.ld_out + p.co_off + nb_st);
            uint4* op2 = reinterpret_cast<uint4*>(out_s2d + f2_st);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              uint32_t ww[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                __nv_bfloat162 h2 = __floats2bfloat162_rn(x[j4 * 8 + e * 2], x[j4 * 8 + e * 2 + 1]);
                ww[e] = *reinterpret_cast<uint32_t*>(&h2);
              }
              op[j4] = make_uint4(ww[0], ww[1], ww[2], ww[3]);
              if (p.s2d_Hp) op2[j4] = make_uint4(ww[0], ww[1], ww[2], ww[3]);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR)
            mbar_arrive_cluster(leader_tmem_empty + 8u * acc);
          else
            mbar_arrive(&tmem_empty[acc]);
        }
      }
    }
    if (tl != nullptr && threadIdx.x == 64) tl[12] = clock64();
  }
  tc_fence_before();
  if (PAIR)
    cluster_sync_all();        // the leader's MMAs read the peer's shared memory and both epilogues arrive on the leader's barriers
  else
    __syncthreads();
  if (warp == 1) {
    if (PAIR)
      tmem_dealloc_pair(tmem_base, 512);
    else
      tmem_dealloc(tmem_base, 512);
  }
  if (tl != nullptr && threadIdx.x == 32) {
    tl[13] = clock64();
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    tl[15] = (long long)gt;
  }
}

int encode_map_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box);   // conv_tc.cu

int init_conv_flat_attrs() {
  cudaError_t e = cudaFuncSetAttribute(conv_flat_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_flat_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(conv_flat_kernel) failed: %s", cudaGetErrorString(e));
    return YAD_ERR_CUDA;
  }
  return YAD_OK;
}

}  // namespace yad

namespace yad {

static long long* g_flat_tlog = nullptr;     // yad_conv_flat_set_timeline

// common launcher: the caller has filled the geometry / epilogue fields and the step lists (grouped by chunk)
static int launch_flat(FlatParams& p, int n_chunks_distinct, int max_off, const void* in, int cin_total, int ld_in,
                       const void* weight, int64_t k_total, int cout_pad, const float* bias, const void* residual, void* out,
                       yad_stream_t stream, void* out_s2d = nullptr, const void* in2 = nullptr, int cin2_total = 0, int ld_in2 = 0) {
  for (int i = 0; i < p.n_steps; ++i)
    p.step_mma[i] = (uint32_t)((p.step_off[i] - p.min_off) * 8) | (p.step_first[i] ? 1u << 30 : 0u) | (p.step_last[i] ? 1u << 31 : 0u);
  YAD_CHECK_ARG((max_off - p.min_off) * 8 < 65536, "yad_conv_flat: filter reach too large");
  p.BN = cout_pad % 128 == 0 ? 128 : 64;
  p.n_ntiles = cout_pad / p.BN;
  p.MT = FL_ACC_COLS / p.BN;              // 2 streams x 2 accumulator stages x MT x BN = 512 TMEM columns
  // CTA pairs (cluster of 2, tcgen05 cta_group::2): every MMA covers 128 rows of each CTA and takes half of the weight block's
  // rows from each CTA's shared memory - half the B-operand reads and half the weight fill per SM (YAD_FLAT_PAIR=0: off)
  static const bool pair_enabled = [] {
    const char* e = getenv("YAD_FLAT_PAIR");
    return !(e && e[0] == '0');
  }();
  const int nsm = sm_count() > 0 ? sm_count() : 148;
  // N = 64 stays single-CTA: measured (layer1 shape) 75 cycles per cta_group::2 MMA against 69 per cta_group::1 MMA - the A
  // operand (4 KB per MMA either way) dominates there and the pair instruction costs more than the 1 KB of B it saves
  const bool pair = pair_enabled && nsm >= 2 && p.BN == 128;
  const int rows_per_super = 128 * p.MT;
  const int patch_rows_raw = rows_per_super + max_off - p.min_off;
  p.patch_rows = (patch_rows_raw + FL_BOX_ROWS - 1) / FL_BOX_ROWS * FL_BOX_ROWS;
  p.patch_bytes = p.patch_rows * 128;
  p.w_bytes = (pair ? p.BN / 2 : p.BN) * 128;          // per CTA
  const int64_t n_mtiles = (p.F + rows_per_super - 1) / rows_per_super;
  const int tiles_per_pair = pair ? 4 : 2;             // two streams (x two CTAs)
  p.n_super = (int)((n_mtiles + tiles_per_pair - 1) / tiles_per_pair * p.n_ntiles);     // tile pairs
  static const int dbg_flags = [] {           // timing experiments only (results are wrong with any bit set)
    const char* e = getenv("YAD_FLAT_DBG");
    return e ? atoi(e) : 0;
  }();
  p.flags = dbg_flags;
  p.tlog = g_flat_tlog;
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)((pair ? 256 : 128) >> 4) << 24);
  // epilogue through the TMA unit: whole 64-channel blocks, one geometry for input, output and residual
  static const bool tma_epi_enabled = [] {
    const char* e = getenv("YAD_FLAT_TMA_EPI");
    return !(e && e[0] == '0');
  }();
  p.tma_epi = (tma_epi_enabled && !p.remap && p.Cout == cout_pad && cout_pad % 64 == 0 && p.co_off % 8 == 0) ? 1 : 0;
  // shared-memory budget: barriers + bias + alignment slack, the epilogue staging buffers, then the two rings
  size_t fixed = 0;
  for (;;) {
    fixed = 1024 + (4 * FL_MAX_NA + 2 * FL_MAX_RING + 8 + 2 * FL_NSTAGE) * 8 + 16 + (size_t)cout_pad * 4 + 64 +
            (p.tma_epi ? (size_t)FL_NSTAGE * FL_STAGE_BYTES : 0);
    const size_t budget = 227 * 1024 - fixed;
    // patch slots (per stream): the patch producer requests a patch the moment a slot is released, i.e. (NA - 1) chunk durations
    // before its first MMA.  Two slots are enough when one chunk's MMAs (of both streams: they share the tensor pipe) outlast the
    // load of the next patch (~4 k cycles for 20-36 KB); everything else goes to the weight ring, which has to cover the L2
    // latency of its 8 / 16 KB blocks.
    static const int na_forced = [] {
      const char* e = getenv("YAD_FLAT_NA");
      return e ? atoi(e) : 0;
    }();
    const int chunk_cycles = 2 * (p.n_steps / (n_chunks_distinct > 0 ? n_chunks_distinct : 1)) * p.MT * 4 * (p.BN == 128 ? 64 : 48);
    p.NA = chunk_cycles >= 4000 ? 2 : 3;
    if (na_forced >= 2 && na_forced <= FL_MAX_NA) p.NA = na_forced;
    while (p.NA > 1 && 2 * (size_t)p.NA * p.patch_bytes + 3 * (size_t)p.w_bytes > budget) --p.NA;
    const bool fits = 2 * (size_t)p.NA * p.patch_bytes + 3 * (size_t)p.w_bytes <= budget;
    if ((!fits || p.NA < 2) && p.tma_epi) {     // the staging buffers cost the patch double-buffering: keep the register-sourced epilogue
      p.tma_epi = 0;
      continue;
    }
    YAD_CHECK_ARG(fits, "yad_conv_flat: patch of %d rows does not fit", p.patch_rows);
    p.NW = (int)((budget - 2 * (size_t)p.NA * p.patch_bytes) / p.w_bytes);
    if (p.NW > FL_MAX_RING) p.NW = FL_MAX_RING;
    break;
  }
  const size_t smem = fixed + 2 * (size_t)p.NA * p.patch_bytes + (size_t)p.NW * p.w_bytes;

  CUtensorMap map_a, map_w;
  {
    const uint64_t dims[2] = {(uint64_t)cin_total, (uint64_t)p.F};
    const uint64_t strides[1] = {(uint64_t)ld_in * 2};
    const uint32_t box[2] = {64u, (uint32_t)FL_BOX_ROWS};
    int rc = encode_map_bf16(&map_a, in, 2, dims, strides, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)k_total, (uint64_t)cout_pad};
    const uint64_t strides[1] = {(uint64_t)k_total * 2};
    const uint32_t box[2] = {64u, (uint32_t)(pair ? p.BN / 2 : p.BN)};
    int rc = encode_map_bf16(&map_w, weight, 2, dims, strides, box);
    if (rc) return rc;
  }
  CUtensorMap map_a2 = map_a;
  if (in2 != nullptr) {
    const uint64_t dims[2] = {(uint64_t)cin2_total, (uint64_t)p.F};
    const uint64_t strides[1] = {(uint64_t)ld_in2 * 2};
    const uint32_t box[2] = {64u, (uint32_t)FL_BOX_ROWS};
    int rc = encode_map_bf16(&map_a2, in2, 2, dims, strides, box);
    if (rc) return rc;
  }
  CUtensorMap map_out = map_a, map_res = map_a;
  if (p.tma_epi) {
    const uint64_t rows_out = (uint64_t)p.F;       // same geometry as the input (no remap)
    const uint32_t box[2] = {64u, 128u};
    {
      const uint64_t dims[2] = {(uint64_t)p.ld_out, rows_out};
      const uint64_t strides[1] = {(uint64_t)p.ld_out * 2};
      int rc = encode_map_bf16(&map_out, out, 2, dims, strides, box);
      if (rc) return rc;
    }
    if (residual != nullptr) {
      const uint64_t dims[2] = {(uint64_t)p.ld_res, rows_out};
      const uint64_t strides[1] = {(uint64_t)p.ld_res * 2};
      int rc = encode_map_bf16(&map_res, residual, 2, dims, strides, box);
      if (rc) return rc;
    }
  }
  const int units = pair ? nsm / 2 : nsm;
  const int n_units = p.n_super < units ? p.n_super : units;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(pair ? 2 * n_units : n_units));
  cfg.blockDim = dim3(FL_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  {    // programmatic dependent launch: same policy as launch_pdl (common.cuh)
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    int mode = pdl_mode();
    if (mode == 1 && cudaStreamIsCapturing(cfg.stream, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone) mode = 0;
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = mode ? 1 : 0;
    ++na;
  }
  if (pair) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  const __nv_bfloat16* res_p = reinterpret_cast<const __nv_bfloat16*>(residual);
  __nv_bfloat16* out_p = reinterpret_cast<__nv_bfloat16*>(out);
  __nv_bfloat16* s2d_p = reinterpret_cast<__nv_bfloat16*>(out_s2d);
  if (pair)
    YAD_CUDA(cudaLaunchKernelEx(&cfg, conv_flat_kernel<true>, map_a, map_a2, map_w, map_out, map_res, p, bias, res_p, out_p, s2d_p));
  else
    YAD_CUDA(cudaLaunchKernelEx(&cfg, conv_flat_kernel<false>, map_a, map_a2, map_w, map_out, map_res, p, bias, res_p, out_p, s2d_p));
  return YAD_OK;
}

static int check_flat_common(const yad_flat_desc* d, const void* in, const void* weight, int cout_pad, const float* bias,
                             const void* residual, const void* out, FlatParams& p) {
  YAD_CHECK_ARG(d && in && weight && bias && out, "yad_conv_flat: null pointer");
  YAD_CHECK_ARG(d->B >= 1 && d->H >= 1 && d->W >= 1 && d->Hp >= d->H && d->Wp >= d->W, "yad_conv_flat: bad geometry");
  YAD_CHECK_ARG(d->Cin % 64 == 0 && d->Cin >= 64, "yad_conv_flat: Cin=%d must be a multiple of 64", d->Cin);
  YAD_CHECK_ARG(d->ld_in % 8 == 0 && d->ld_in >= d->Cin, "yad_conv_flat: bad ld_in=%d", d->ld_in);
  YAD_CHECK_ARG(cout_pad % 64 == 0 && cout_pad >= d->Cout && d->Cout % 32 == 0,
                "yad_conv_flat: Cout=%d must be a multiple of 32 and cout_pad=%d a multiple of 64", d->Cout, cout_pad);
  YAD_CHECK_ARG(d->ld_out % 8 == 0 && d->co_off % 8 == 0 && d->ld_out >= d->co_off + d->Cout, "yad_conv_flat: bad ld_out/co_off");
  YAD_CHECK_ARG(residual == nullptr || (d->ld_res % 8 == 0 && d->ld_res >= d->Cout), "yad_conv_flat: bad ld_res");
  YAD_CHECK_ARG((reinterpret_cast<uintptr_t>(in) % 16 == 0) && (reinterpret_cast<uintptr_t>(weight) % 16 == 0) &&
                    (reinterpret_cast<uintptr_t>(out) % 16 == 0) && (reinterpret_cast<uintptr_t>(residual) % 16 == 0),
                "yad_conv_flat: pointers must be 16-byte aligned");
  const int Hpo = d->Hp_out > 0 ? d->Hp_out : d->Hp, Wpo = d->Wp_out > 0 ? d->Wp_out : d->Wp;
  YAD_CHECK_ARG(Hpo >= d->H && Wpo >= d->W, "yad_conv_flat: output pitches smaller than the image");
  memset(&p, 0, sizeof(p));
  p.H = d->H;
  p.W = d->W;
  p.Hp = d->Hp;
  p.Wp = d->Wp;
  p.Hpo = Hpo;
  p.Wpo = Wpo;
  p.remap = (Hpo != d->Hp || Wpo != d->Wp) ? 1 : 0;
  {
    auto magic = [](uint32_t dv, uint32_t& mul, uint32_t& sh) {
      uint32_t sft = 0;
      while ((1u << sft) < dv) ++sft;
      sh = 31 + sft;
      mul = (uint32_t)((((uint64_t)1 << sh) + dv - 1) / dv);
    };
    magic((uint32_t)d->Hp, p.hp_mul, p.hp_sh);
    magic((uint32_t)d->Wp, p.wp_mul, p.wp_sh);
  }
  p.F = (int64_t)d->B * d->Wp * d->Hp;
  YAD_CHECK_ARG(p.F < (int64_t)1 << 31 && (int64_t)d->B * Wpo * Hpo < (int64_t)1 << 31, "yad_conv_flat: tensor too large");
  p.Cout = d->Cout;
  p.ld_out = d->ld_out;
  p.co_off = d->co_off;
  p.ld_res = d->ld_res;
  p.act = d->act;
  return YAD_OK;
}

}  // namespace yad

static int conv_flat_impl(const yad_flat_desc* d, const void* in, const void* weight, int32_t cout_pad, const float* bias,
                          const void* residual, void* out, void* out_s2d, int32_t Hp2, int32_t Wp2, yad_stream_t stream) {
  using namespace yad;
  FlatParams p;
  int rc = check_flat_common(d, in, weight, cout_pad, bias, residual, out, p);
  if (rc) return rc;
  if (out_s2d != nullptr) {
    YAD_CHECK_ARG(Hp2 >= (d->H + 1) / 2 && Wp2 >= (d->W + 1) / 2 && reinterpret_cast<uintptr_t>(out_s2d) % 16 == 0 && d->Cout % 32 == 0 &&
                      (int64_t)d->B * Wp2 * Hp2 * 4 * d->Cout < ((int64_t)1 << 32),
                  "yad_conv_flat_s2d: space-to-depth pitches (%d,%d) too small for a %dx%d image, or tensor too large", Hp2, Wp2, d->H, d->W);
    p.s2d_Hp = Hp2;
    p.s2d_Wp = Wp2;
  }
  YAD_CHECK_ARG(d->kh >= 1 && d->kw >= 1 && d->kh * d->kw <= 49 && d->ph >= 0 && d->pw >= 0 && d->ph < d->kh && d->pw < d->kw,
                "yad_conv_flat: bad kernel/padding");
  YAD_CHECK_ARG(d->kh - 1 - d->ph == d->ph && d->kw - 1 - d->pw == d->pw, "yad_conv_flat: only 'same' (output size = input size) convs");
  // taps: skip those that only ever see padding; the rest must be covered by the halo
  int taps_dh[49], taps_dw[49], taps_idx[49], n_taps = 0, reach_h = 0, reach_w = 0;
  for (int kh = 0; kh < d->kh; ++kh) {
    const int dh = kh - d->ph;
    if (dh <= -d->H || dh >= d->H) continue;
    for (int kw = 0; kw < d->kw; ++kw) {
      const int dw = kw - d->pw;
      if (dw <= -d->W || dw >= d->W) continue;
      taps_dh[n_taps] = dh;
      taps_dw[n_taps] = dw;
      taps_idx[n_taps] = kh * d->kw + kw;
      reach_h = dh < 0 ? (-dh > reach_h ? -dh : reach_h) : (dh > reach_h ? dh : reach_h);
      reach_w = dw < 0 ? (-dw > reach_w ? -dw : reach_w) : (dw > reach_w ? dw : reach_w);
      ++n_taps;
    }
  }
  YAD_CHECK_ARG(d->Hp - d->H >= reach_h && d->Wp - d->W >= reach_w,
                "yad_conv_flat: halo (%d,%d) smaller than the filter reach (%d,%d)", d->Hp - d->H, d->Wp - d->W, reach_h, reach_w);
  const int n_chunks = d->Cin / 64;
  YAD_CHECK_ARG(n_taps * n_chunks <= FL_MAX_STEPS, "yad_conv_flat: %d taps x %d chunks exceed %d steps", n_taps, n_chunks, FL_MAX_STEPS);
  int min_off = 0, max_off = 0, ns = 0;
  for (int c = 0; c < n_chunks; ++c) {
    for (int t = 0; t < n_taps; ++t, ++ns) {
      const int off = taps_dw[t] * d->Hp + taps_dh[t];
      min_off = off < min_off ? off : min_off;
      max_off = off > max_off ? off : max_off;
      p.step_off[ns] = (int16_t)off;
      p.step_chunk[ns] = (int16_t)c;
      p.step_first[ns] = (t == 0);
      p.step_last[ns] = (t == n_taps - 1);
      p.step_wk[ns] = taps_idx[t] * d->Cin + c * 64;
    }
  }
  p.n_steps = ns;
  p.min_off = min_off;
  return launch_flat(p, n_chunks, max_off, in, d->Cin, d->ld_in, weight, (int64_t)d->kh * d->kw * d->Cin, cout_pad, bias, residual,
                     out, stream, out_s2d);
}

/* debug: device buffer of 8 int64 per CTA (148 CTAs) that the next conv_flat launches fill with the MMA warp's cycle budget
 * [total, wait accumulator stage, wait patch, wait weights, issue block]; NULL switches it off */
extern "C" int yad_conv_flat_set_timeline(void* dev_buf) {
  yad::g_flat_tlog = reinterpret_cast<long long*>(dev_buf);
  return YAD_OK;
}

extern "C" int yad_conv_flat(const yad_flat_desc* d, const void* in, const void* weight, int32_t cout_pad,
                             const float* bias, const void* residual, void* out, int32_t flags, yad_stream_t stream) {
  using namespace yad;
  YAD_CHECK_ARG(flags == 0, "yad_conv_flat: flags must be 0");
  return conv_flat_impl(d, in, weight, cout_pad, bias, residual, out, nullptr, 0, 0, stream);
}

extern "C" int yad_conv_flat_s2d(const yad_flat_desc* d, const void* in, const void* weight, int32_t cout_pad, const float* bias,
                                 const void* residual, void* out, void* out_s2d, int32_t Hp2, int32_t Wp2, yad_stream_t stream) {
  using namespace yad;
  YAD_CHECK_ARG(out_s2d != nullptr, "yad_conv_flat_s2d: null pointer");
  return conv_flat_impl(d, in, weight, cout_pad, bias, residual, out, out_s2d, Hp2, Wp2, stream);
}

static int conv_flat_taps_impl(const yad_flat_desc* d, int32_t n_steps, const int32_t* src, const int32_t* chunk, const int32_t* dh,
                               const int32_t* dw, const int32_t* wk, int64_t k_total, const void* in, const void* in2, int32_t cin2,
                               int32_t ld_in2, const void* weight, int32_t cout_pad, const float* bias, const void* residual, void* out,
                               yad_stream_t stream) {
  using namespace yad;
  FlatParams p;
  int rc = check_flat_common(d, in, weight, cout_pad, bias, residual, out, p);
  if (rc) return rc;
  YAD_CHECK_ARG(in2 == nullptr || (src != nullptr && cin2 % 64 == 0 && cin2 >= 64 && ld_in2 % 8 == 0 && ld_in2 >= cin2 &&
                                   reinterpret_cast<uintptr_t>(in2) % 16 == 0),
                "yad_conv_flat_taps2: bad second input (Cin2=%d, ld_in2=%d)", cin2, ld_in2);
  YAD_CHECK_ARG(chunk && dh && dw && wk && n_steps >= 1 && n_steps <= FL_MAX_STEPS, "yad_conv_flat_taps: bad step list (1..%d steps)",
                FL_MAX_STEPS);
  int min_off = 0, max_off = 0, n_distinct = 0;
  for (int i = 0; i < n_steps; ++i) {
    const int sr = (src != nullptr && in2 != nullptr) ? src[i] : 0;
    YAD_CHECK_ARG(sr == 0 || sr == 1, "yad_conv_flat_taps2: step %d: source must be 0 or 1", i);
    YAD_CHECK_ARG(chunk[i] >= 0 && chunk[i] * 64 + 64 <= (sr ? cin2 : d->Cin), "yad_conv_flat_taps: step %d reads chunk %d outside Cin=%d", i,
                  chunk[i], sr ? cin2 : d->Cin);
    YAD_CHECK_ARG(i == 0 || (sr ? FL_SRC2 : 0) + chunk[i] >= p.step_chunk[i - 1],
                  "yad_conv_flat_taps: steps must be grouped by (source, ascending chunk)");
    YAD_CHECK_ARG(wk[i] >= 0 && wk[i] % 8 == 0 && (int64_t)wk[i] + 64 <= k_total, "yad_conv_flat_taps: bad weight offset in step %d", i);
    const int adh = dh[i] < 0 ? -dh[i] : dh[i], adw = dw[i] < 0 ? -dw[i] : dw[i];
    YAD_CHECK_ARG(adh <= d->Hp - d->H && adw <= d->Wp - d->W, "yad_conv_flat_taps: step %d reaches (%d,%d) beyond the halo (%d,%d)", i,
                  dh[i], dw[i], d->Hp - d->H, d->Wp - d->W);
    const int off = dw[i] * d->Hp + dh[i];
    min_off = off < min_off ? off : min_off;
    max_off = off > max_off ? off : max_off;
    p.step_off[i] = (int16_t)off;
    p.step_chunk[i] = (int16_t)((sr ? FL_SRC2 : 0) + chunk[i]);
    p.step_wk[i] = wk[i];
  }
  for (int i = 0; i < n_steps; ++i) {
    p.step_first[i] = (i == 0 || p.step_chunk[i] != p.step_chunk[i - 1]);
    p.step_last[i] = (i == n_steps - 1 || p.step_chunk[i + 1] != p.step_chunk[i]);
    if (p.step_first[i]) ++n_distinct;
  }
  p.n_steps = n_steps;
  p.min_off = min_off;
  return launch_flat(p, n_distinct, max_off, in, d->Cin, d->ld_in, weight, k_total, cout_pad, bias, residual, out, stream, nullptr, in2, cin2,
                     ld_in2);
}

extern "C" int yad_conv_flat_taps(const yad_flat_desc* d, int32_t n_steps, const int32_t* chunk, const int32_t* dh,
                                  const int32_t* dw, const int32_t* wk, int64_t k_total, const void* in, const void* weight,
                                  int32_t cout_pad, const float* bias, const void* residual, void* out, yad_stream_t stream) {
  return conv_flat_taps_impl(d, n_steps, nullptr, chunk, dh, dw, wk, k_total, in, nullptr, 0, 0, weight, cout_pad, bias, residual, out, stream);
}

extern "C" int yad_conv_flat_taps2(const yad_flat_desc* d, int32_t n_steps, const int32_t* src, const int32_t* chunk, const int32_t* dh,
                                   const int32_t* dw, const int32_t* wk, int64_t k_total, const void* in, const void* in2, int32_t cin2,
                                   int32_t ld_in2, const void* weight, int32_t cout_pad, const float* bias, const void* residual, void* out,
                                   yad_stream_t stream) {
  using namespace yad;
  YAD_CHECK_ARG(in2 != nullptr && src != nullptr, "yad_conv_flat_taps2: null pointer");
  return conv_flat_taps_impl(d, n_steps, src, chunk, dh, dw, wk, k_total, in, in2, cin2, ld_in2, weight, cout_pad, bias, residual, out, stream);
}
