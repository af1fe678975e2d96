// Implicit-GEMM 3x3(x3) convolution on tcgen05 with a shared-memory HALO tile — persistent, warp-specialised.
//
// Work item ("tile") = a macro tile of 16 rows x (8*mt) columns of one image plane (mt MMA tiles of 16x8 = 128
// pixels) x n_blk output channels. A CTA (one per SM) walks tiles blockIdx.x, +gridDim.x, ...
// Per input-channel chunk the A producer TMA-loads ONE halo tile (ck channels x (8*mt+2) x 18 x kd pixels; out of
// bounds = zero = the convolution's padding) and all 9 / 27 filter taps run out of it: the A operand of tap
// (dz,dy,dx), MMA tile j is the same shared-memory tile addressed at row
//     ((dz*18 + dy) * (8*mt+2) + 8*j + dx)
// with 8-row core-matrix groups (8*mt+2) rows apart (stride-byte-offset of the smem descriptor). The 128B / 64B /
// 32B TMA swizzle is a function of the absolute shared-memory address, so row-shifted starts need no re-layout
// (verified on B200 by tools/tc_probe.cu -> profiles/r01_tc_descriptor_probe.txt).
// Versus one TMA box per tap (conv_tc.cuh): ~9x fewer TMA rows / L2->SM bytes for A, and the mt MMA tiles share
// every weight tile.
// Pipelines: halo ring (a_full/a_empty), weight-tile ring (b_full/b_empty), two TMEM accumulator stages
// (acc_full/acc_empty) so the epilogue of tile i overlaps the MMAs of tile i+1.
// Warp roles: 0 = halo (A) producer, 1 = TMEM alloc + MMA issuer, 2 = weight (B) producer, 4..11 = epilogue (two
// groups of four warps; a warp reads the TMEM lane quarter warp % 4).
#pragma once
#include "conv_tc.cuh"

namespace biu {

constexpr int kHaloRows = 18;      // 16 output rows + 2
constexpr int kMaxBStages = 8;
constexpr int kMaxABufs = 3;
constexpr int kHaloThreads = 384;   // warps: 0 A producer, 1 MMA, 2 B producer, 3 idle, 4..11 epilogue

struct ConvHaloParams {
  int W, H, D, B;
  int mt;                          // MMA tiles (16x8 pixels) per work item along x
  int tiles_x, tiles_y;            // macro tiles per image plane
  int n_blocks;                    // output-channel blocks
  int total_tiles;                 // tiles_x * tiles_y * D * B * n_blocks
  int kd;                          // 1 (2D) or 3
  int cin_chunks, ck, row_bytes;
  int n_blk;
  int a_bufs, b_stages;
  uint32_t a_buf_bytes, b_stage_bytes;
  int mode;
  float slope;
  const float* scale;
  const float* shift;
  void* out;
  int out_ctot, out_coff;
  int head_n;
  const float* head_w;
  const float* head_b;
  int head_act[kMaxHead];
  float* out_val;
  uint8_t* out_u8;
};

struct HaloTile { int x0, y0, z0, b0, n0; };

__device__ __forceinline__ HaloTile halo_decode(const ConvHaloParams& p, int t) {
  HaloTile r;
  const int tx = t % p.tiles_x; t /= p.tiles_x;
  const int ty = t % p.tiles_y; t /= p.tiles_y;
  r.z0 = t % p.D; t /= p.D;
  r.b0 = t % p.B; t /= p.B;
  r.n0 = t * p.n_blk;
  r.x0 = tx * 8 * p.mt;
  r.y0 = ty * 16;
  return r;
}

template <int ESZ, int KS>
__global__ void __launch_bounds__(kHaloThreads, 1) conv_halo_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                    const __grid_constant__ CUtensorMap tmB,
                                                                    const ConvHaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t a_full[kMaxABufs], a_empty[kMaxABufs];
  __shared__ uint64_t b_full[kMaxBStages], b_empty[kMaxBStages];
  __shared__ uint64_t acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_base = smem_base + p.a_bufs * p.a_buf_bytes;
  const int pw = 8 * p.mt + 2;                                   // halo tile pitch in pixels
  const uint32_t rb = p.row_bytes;
  const int taps = 9 * p.kd;
  const uint32_t acc_cols = (uint32_t)(p.mt * p.n_blk);          // columns of one accumulator stage

  uint32_t ncols = 32;
  while (ncols < 2 * acc_cols) ncols <<= 1;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < p.a_bufs; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int s = 0; s < p.b_stages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_slot, ncols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ================================== halo (A) producer ===================================
    if (elect_one()) {
      const uint32_t halo_tx = (uint32_t)(p.kd * kHaloRows * pw) * rb;
      int ab = 0;
      uint32_t aph = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const HaloTile tl = halo_decode(p, t);
        for (int ch = 0; ch < p.cin_chunks; ++ch) {
          mbar_wait(&a_empty[ab], aph ^ 1, 0x400 + ab);
          mbar_arrive_expect_tx(&a_full[ab], halo_tx);
          // The halo is fetched as 2-row boxes issued back to back: one TMA operation keeps only a few dozen L2
          // requests in flight, many concurrent ones are needed to cover the L2 / HBM latency.
          const uint32_t box_bytes = 2u * (uint32_t)pw * rb;
          uint32_t dst = smem_base + ab * p.a_buf_bytes;
          for (int dz = 0; dz < p.kd; ++dz)
            for (int r2 = 0; r2 < kHaloRows / 2; ++r2, dst += box_bytes)
              asm volatile(
                  "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
                  "%5, %6, %7}], [%2];" ::"r"(dst),
                  "l"(reinterpret_cast<uint64_t>(&tmA)), "r"(smem_u32(&a_full[ab])), "r"(ch * p.ck), "r"(tl.x0 - 1),
                  "r"(tl.y0 - 1 + 2 * r2), "r"(tl.z0 - (p.kd >> 1) + dz), "r"(tl.b0)
                  : "memory");
          if (++ab == p.a_bufs) { ab = 0; aph ^= 1; }
        }
      }
    }
  } else if (warp == 2) {
    // ================================= weight (B) producer ==================================
    if (elect_one()) {
      const uint32_t b_tx = (uint32_t)p.n_blk * rb;
      int s = 0;
      uint32_t bph = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const HaloTile tl = halo_decode(p, t);
        for (int ch = 0; ch < p.cin_chunks; ++ch)
          for (int tap = 0; tap < taps; ++tap) {
            mbar_wait(&b_empty[s], bph ^ 1, 0x500 + s);
            mbar_arrive_expect_tx(&b_full[s], b_tx);
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
                "%5}], [%2];" ::"r"(b_base + s * p.b_stage_bytes),
                "l"(reinterpret_cast<uint64_t>(&tmB)), "r"(smem_u32(&b_full[s])), "r"(ch * p.ck), "r"(tl.n0),
                "r"(tap)
                : "memory");
            if (++s == p.b_stages) { s = 0; bph ^= 1; }
          }
      }
    }
  } else if (warp == 1) {
    // ====================================== MMA issuer ======================================
    if (elect_one()) {
      const uint32_t layout = rb == 128 ? 2u : (rb == 64 ? 4u : 6u);
      const uint32_t idesc = make_idesc(ESZ == 2 ? 1u : 2u, (uint32_t)p.n_blk);
      // Descriptors: only the 14-bit start-address field (low word, address >> 4) changes between MMAs and it
      // never carries out of the field (shared memory < 256 KB), so they are advanced with 32-bit adds.
      const uint64_t a_tmpl = make_smem_desc(0, (uint32_t)pw * rb, layout);
      const uint64_t b_tmpl = make_smem_desc(0, 8u * rb, layout);
      const uint32_t a_hi = (uint32_t)(a_tmpl >> 32), b_hi = (uint32_t)(b_tmpl >> 32);
      const uint32_t a_lo0 = (uint32_t)a_tmpl + ((smem_base & 0x3FFFF) >> 4);
      const uint32_t b_lo0 = (uint32_t)b_tmpl + ((b_base & 0x3FFFF) >> 4);
      const uint32_t j_step = (8u * rb) >> 4;        // next MMA tile: 8 pixels further
      const uint32_t px_step = rb >> 4;              // one pixel
      const uint32_t row_step = ((uint32_t)pw * rb) >> 4;
      const uint32_t abuf_step = p.a_buf_bytes >> 4, bst_step = p.b_stage_bytes >> 4;
      int ab = 0, bs = 0, it = 0;
      uint32_t aph = 0, bph = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
        const int as = it & 1;
        mbar_wait(&acc_empty[as], ((it >> 1) & 1) ^ 1, 0x900 + as);     // epilogue has drained this stage
        tc_fence_after();
        const uint32_t tacc = tmem_base + as * acc_cols;
        uint32_t accum = 0;                                             // first k-block of a tile overwrites
        for (int ch = 0; ch < p.cin_chunks; ++ch) {
          mbar_wait(&a_full[ab], aph, 0x600 + ab);
          const uint32_t a_buf_lo = a_lo0 + ab * abuf_step;
          for (int r = 0; r < 3 * p.kd; ++r) {                          // r = dz*3 + dy
            const int dz = r / 3, dy = r - 3 * dz;
            const uint32_t a_row_lo = a_buf_lo + (uint32_t)(dz * kHaloRows + dy) * row_step;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              mbar_wait(&b_full[bs], bph, 0x700 + bs);
              tc_fence_after();
              const uint32_t b_lo = b_lo0 + bs * bst_step;
              uint32_t a_lo = a_row_lo + dx * px_step;
              uint32_t tcol = tacc;
              for (int j = 0; j < p.mt; ++j, a_lo += j_step, tcol += p.n_blk) {
#pragma unroll
                for (int k = 0; k < KS; ++k) {
                  const uint64_t ad = ((uint64_t)a_hi << 32) | (a_lo + 2 * k);   // +32 bytes per k-step
                  const uint64_t bd = ((uint64_t)b_hi << 32) | (b_lo + 2 * k);
                  const uint32_t acc = k == 0 ? accum : 1u;
                  if (ESZ == 2) tc_mma_f16(tcol, ad, bd, idesc, acc); else tc_mma_tf32(tcol, ad, bd, idesc, acc);
                }
              }
              accum = 1;
              tc_commit(&b_empty[bs]);
              if (++bs == p.b_stages) { bs = 0; bph ^= 1; }
            }
          }
          tc_commit(&a_empty[ab]);
          if (++ab == p.a_bufs) { ab = 0; aph ^= 1; }
        }
        tc_commit(&acc_full[as]);
      }
    }
  } else if (warp >= 4) {
    // ======================================= epilogue =======================================
    // Work units = (MMA tile j, column chunk); the two warp groups take alternate units. With the fused head a
    // thread needs the whole channel row of its pixel, so there the unit is the MMA tile.
    const int grp = warp & 3;                      // TMEM lane quarter
    const int egrp = (warp - 4) >> 2;              // epilogue group 0 / 1
    const int m = grp * 32 + lane;                 // accumulator row: pixel (m >> 3, m & 7) of each MMA tile
    const int cw = (p.n_blk % 32 == 0) ? 32 : 16;  // chunk width
    const int nchunks = p.n_blk / cw;
    const bool head = p.mode == EPI_HEAD;
    int it = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
      const HaloTile tl = halo_decode(p, t);
      const int as = it & 1;
      mbar_wait(&acc_full[as], (it >> 1) & 1, 0x800 + as);
      tc_fence_after();
      const int py = tl.y0 + (m >> 3);
      const uint32_t trow = tmem_base + as * acc_cols + ((uint32_t)(grp * 32) << 16);
      const long long plane_pix = ((long long)tl.b0 * p.D + tl.z0) * p.H;
      for (int j = 0; j < p.mt; ++j) {
        if (head && (j & 1) != egrp) continue;
        const int px = tl.x0 + 8 * j + (m & 7);
        const bool valid = px < p.W && py < p.H;
        const long long pix = (plane_pix + py) * p.W + px;
        float hacc[kMaxHead];
#pragma unroll
        for (int h = 0; h < kMaxHead; ++h) hacc[h] = 0.f;
        for (int c = 0; c < nchunks; ++c) {
          if (!head && ((j * nchunks + c) & 1) != egrp) continue;
          const int c0 = c * cw;
          const int n = tl.n0 + c0;
          uint32_t acc[32];
          if (cw == 32) {
            tmem_ld32(trow + j * p.n_blk + c0, acc);
          } else {
            uint32_t lo[16];
            tmem_ld16(trow + j * p.n_blk + c0, lo);
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = lo[i];
          }
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            if (q * 16 >= cw) break;
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float a = fmaf(__uint_as_float(acc[q * 16 + i]), __ldg(p.scale + n + q * 16 + i),
                             __ldg(p.shift + n + q * 16 + i));
              v[i] = a > 0.f ? a : a * p.slope;
            }
            if (head) {
              for (int h = 0; h < p.head_n; ++h) {
                float sacc = hacc[h];
#pragma unroll
                for (int i = 0; i < 16; ++i) sacc = fmaf(v[i], __ldg(p.head_w + h * p.n_blk + c0 + q * 16 + i), sacc);
                hacc[h] = sacc;
              }
            }
            if (valid && p.out != nullptr) {
              const long long off = pix * p.out_ctot + p.out_coff + n + q * 16;
              if (ESZ == 2) {
                uint32_t w[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  __nv_bfloat162 b2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                  w[i] = *reinterpret_cast<uint32_t*>(&b2);
                }
                uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + off);
                dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
                dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
              } else {
                float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + off);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  dst[i] = make_float4(round_tf32(v[4 * i]), round_tf32(v[4 * i + 1]), round_tf32(v[4 * i + 2]),
                                       round_tf32(v[4 * i + 3]));
              }
            }
          }
        }
        if (head && valid) {
          const long long plane = (long long)p.D * p.H * p.W;
          const long long sp = ((long long)tl.z0 * p.H + py) * p.W + px;
          for (int h = 0; h < p.head_n; ++h) {
            const float val = apply_head_act(hacc[h] + __ldg(p.head_b + h), p.head_act[h]);
            const long long o = ((long long)tl.b0 * p.head_n + h) * plane + sp;
            if (p.out_val) p.out_val[o] = val;
            if (p.out_u8) p.out_u8[o] = (uint8_t)(val * 255.0f);
          }
        }
      }
      // this warp is done reading the accumulator stage: hand it back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, ncols);
}

}  // namespace biu
