// Row-streaming 3x3(x3) convolution for NARROW layers (Cout <= 32) on tcgen05: the dy taps are folded into N.
//
// Why: with M = 128 the tensor core fetches its operands from shared memory at ~85 B/clk (measured,
// tools/conv_bench.cu): an MMA of N = 32 costs ~60 cycles for 16 cycles of math, so the 9-tap / 27-tap form of
// conv_halo.cuh is operand-fetch bound at ~25 % of the tensor peak for the full-resolution layers. Here one MMA
// multiplies an input row segment with the weights of all three dy taps at once (N = 3 * Cout):
//     E[r][x][dy, co] = sum_{dz, dx, ci} in[z + dz - 1][r][x + dx - 1][ci] * W[co][ci][dz][dy][dx]
//     out[y][x][co]   = E[y - 1][x][0, co] + E[y][x][1, co] + E[y + 1][x][2, co]
// i.e. 3 (x kd) MMAs of N = 96 per 16 input channels instead of 9 (x kd) of N = 32: ~2.1x fewer operand bytes per
// useful MAC. The epilogue adds the three partial rows while it drains TMEM.
//
// Work item = a strip of 128 pixels (one MMA M tile = 128 consecutive x of one image row) x RB output rows of one
// plane. Input rows y0-1 .. y0+RB stream through a shared-memory ring (one TMA box of 130 pixels per row, plane and
// channel chunk; out-of-image rows / columns / planes are zero-filled = the convolution's padding). Every input
// row owns a TMEM slot of 3*Cout fp32 columns (ring of 512 / (3*Cout) slots); output row y is ready once input row
// y+1 has been accumulated. The folded weights of the layer stay resident in shared memory.
// Warp roles: 0 = row (A) producer, 1 = TMEM alloc + MMA issuer (warp-uniform, see conv_halo.cuh), 2 = weight
// loader, 4..11 = epilogue: warp w drains TMEM lane quarter w % 4 (32 pixels) and channel half (w - 4) / 4, all
// eight warps work on the same output row, so a slot is simply released by eight arrivals after the row that last
// read it.
#pragma once
#include "conv_halo.cuh"

namespace biu {

constexpr int kRowsMaxASlots = 24;
constexpr int kRowsMaxTSlots = 10;
constexpr int kRowsThreads = 384;
constexpr int kRowsPx = 130;                 // 128 pixels + 1 halo column on each side

struct ConvRowsParams {
  int W, H, D, B;
  int strips, rblocks, RB;       // strips of 128 px per row, row blocks per plane, rows per block
  int total_items;               // strips * rblocks * D * B
  int kd;                        // 1 (2D) or 3
  int cin_chunks, ck, row_bytes;
  int cp;                        // padded output channels (16 or 32); N of the folded MMA = 3 * cp
  int a_slots;                   // shared-memory ring: one slot = one input row of one plane, all channel chunks
  uint32_t a_slot_bytes, a_chunk_bytes;
  uint32_t w_tile_bytes;         // one folded weight tile [(dy, co)][ck]; kd * 3 * cin_chunks of them
  int t_slots;                   // TMEM ring: slots of 3 * cp columns
  int mode;                      // EPI_CONV or EPI_HEAD
  float slope;
  const float* scale;
  const float* shift;
  void* out;
  int out_ctot, out_coff;
  void* pool_out;                // EPI_CONV, 2D: fused MaxPool2d(2) (may be null)
  int pool_ctot, pool_coff;
  int head_n;
  const float* head_w;           // [head_n][cp]
  const float* head_b;
  int head_act[kMaxHead];
  float* out_val;
  uint8_t* out_u8;
};

template <int HC>
__device__ __forceinline__ void tmem_ld_hc(uint32_t taddr, uint32_t (&r)[16]) {
  if (HC == 16) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
        "%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
  } else {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
  }
}
template <int HC>
__device__ __forceinline__ void tmem_wait3(uint32_t (&a)[16], uint32_t (&b)[16], uint32_t (&c)[16]) {
  if (HC == 16) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]),
                   "+r"(a[8]), "+r"(a[9]), "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15]),
                   "+r"(b[0]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7]),
                   "+r"(b[8]), "+r"(b[9]), "+r"(b[10]), "+r"(b[11]), "+r"(b[12]), "+r"(b[13]), "+r"(b[14]), "+r"(b[15]),
                   "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]), "+r"(c[4]), "+r"(c[5]), "+r"(c[6]), "+r"(c[7]),
                   "+r"(c[8]), "+r"(c[9]), "+r"(c[10]), "+r"(c[11]), "+r"(c[12]), "+r"(c[13]), "+r"(c[14]), "+r"(c[15])
                 :
                 : "memory");
  } else {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]),
                   "+r"(b[0]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7]),
                   "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]), "+r"(c[4]), "+r"(c[5]), "+r"(c[6]), "+r"(c[7])
                 :
                 : "memory");
  }
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

struct RowsItem { int x0, y0, rows, z, b; };

__device__ __forceinline__ RowsItem rows_decode(const ConvRowsParams& p, int t) {
  RowsItem r;
  const int sx = t % p.strips; t /= p.strips;
  const int rb = t % p.rblocks; t /= p.rblocks;
  r.z = t % p.D; t /= p.D;
  r.b = t;
  r.x0 = sx * 128;
  r.y0 = rb * p.RB;
  r.rows = min(p.RB, p.H - r.y0);
  return r;
}

// Epilogue of all work items for one warp. HC = channels per warp (cp / 2).
template <int ESZ, int HC, int MODE, bool POOL>
__device__ __forceinline__ void rows_epilogue(const ConvRowsParams& p, uint32_t tmem_base, uint64_t* t_full,
                                              uint64_t* t_empty, const float* s_scale, const float* s_shift,
                                              const float* s_headw, float* s_part, int warp, int lane) {
  const int q = warp & 3;                        // TMEM lane quarter: pixels 32q .. 32q+31 of the strip
  const int hsel = (warp - 4) >> 2;              // channel half
  const int c0 = hsel * HC;                      // first channel of this warp
  const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
  const int nfold = 3 * p.cp;
  int ts = 0;                                    // ring position (slot) and use parity of the item's first input row
  uint32_t tph = 0;
  float sc[HC], sh[HC];
#pragma unroll
  for (int i = 0; i < HC; ++i) { sc[i] = s_scale[c0 + i]; sh[i] = s_shift[c0 + i]; }
  uint32_t carry[HC / 2 * (ESZ == 2 ? 1 : 2)];   // POOL: previous (even) row, already max-ed over the x pair
  (void)carry;

  for (int t = blockIdx.x; t < p.total_items; t += gridDim.x) {
    const RowsItem it = rows_decode(p, t);
    const int px = it.x0 + q * 32 + lane;
    const bool col_ok = px < p.W;
    const long long plane_row0 = ((long long)it.b * p.D + it.z) * p.H + it.y0;
    char* out_px = nullptr;
    if (MODE == EPI_CONV)
      out_px = reinterpret_cast<char*>(p.out) + ((plane_row0 * p.W + px) * p.out_ctot + p.out_coff + c0) * ESZ;
    const long long out_row_bytes = (long long)p.W * p.out_ctot * ESZ;
    char* pool_px = nullptr;
    long long pool_row_bytes = 0;
    if (POOL) {
      const long long prow0 = ((long long)it.b * p.D + it.z) * (p.H >> 1) + (it.y0 >> 1);
      pool_px = reinterpret_cast<char*>(p.pool_out) + ((prow0 * (p.W >> 1) + (px >> 1)) * p.pool_ctot + p.pool_coff + c0) * ESZ;
      pool_row_bytes = (long long)(p.W >> 1) * p.pool_ctot * ESZ;
    }
    // slots / parities of input rows o, o+1, o+2 (relative to the item's first input row y0-1)
    int s0 = ts; uint32_t ph0 = tph;
    int s1 = s0 + 1; uint32_t ph1 = ph0; if (s1 == p.t_slots) { s1 = 0; ph1 ^= 1; }
    int s2 = s1 + 1; uint32_t ph2 = ph1; if (s2 == p.t_slots) { s2 = 0; ph2 ^= 1; }
    for (int o = 0; o < it.rows; ++o) {
      mbar_wait(&t_full[s2], ph2, 0xA00 + s2);   // rows complete in order: o and o+1 are done as well
      tc_fence_after();
      uint32_t e0[16], e1[16], e2[16];
      tmem_ld_hc<HC>(tmem_base + lane_addr + (uint32_t)(s0 * nfold + c0), e0);
      tmem_ld_hc<HC>(tmem_base + lane_addr + (uint32_t)(s1 * nfold + p.cp + c0), e1);
      tmem_ld_hc<HC>(tmem_base + lane_addr + (uint32_t)(s2 * nfold + 2 * p.cp + c0), e2);
      tmem_wait3<HC>(e0, e1, e2);
      float v[HC];
#pragma unroll
      for (int i = 0; i < HC; ++i) {
        float a = (__uint_as_float(e0[i]) + __uint_as_float(e1[i])) + __uint_as_float(e2[i]);
        a = fmaf(a, sc[i], sh[i]);
        a = fmaxf(a, a * p.slope);               // LeakyReLU for 0 <= slope <= 1
        if (ESZ == 4) a = round_tf32(a);
        v[i] = a;
      }
      // slot of input row o is not needed any more (rows o+1, o+2 are still read by the next output rows)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[s0]);

      if (MODE == EPI_HEAD) {
        float part[kMaxHead];
#pragma unroll
        for (int h = 0; h < kMaxHead; ++h) {
          if (h >= p.head_n) break;
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < HC; ++i) s = fmaf(v[i], s_headw[h * p.cp + c0 + i], s);
          part[h] = s;
        }
        // channel halves are combined through shared memory (double-buffered by row parity, one named barrier of
        // the two warps that share a lane quarter per row)
        float* buf = s_part + ((o & 1) * 128 + q * 32 + lane) * kMaxHead;
        if (hsel == 1) {
#pragma unroll
          for (int h = 0; h < kMaxHead; ++h) { if (h >= p.head_n) break; buf[h] = part[h]; }
        }
        named_bar_sync(1 + q, 64);
        if (hsel == 0 && col_ok) {
          const long long plane = (long long)p.D * p.H * p.W;
          const long long sp = ((long long)it.z * p.H + it.y0 + o) * p.W + px;
#pragma unroll
          for (int h = 0; h < kMaxHead; ++h) {
            if (h >= p.head_n) break;
            const float val = apply_head_act(part[h] + buf[h] + __ldg(p.head_b + h), p.head_act[h]);
            const long long o2 = ((long long)it.b * p.head_n + h) * plane + sp;
            if (p.out_val) p.out_val[o2] = val;
            if (p.out_u8) p.out_u8[o2] = (uint8_t)(val * 255.0f);     // unet/predict.py:200 truncating cast
          }
        }
      } else if (ESZ == 2) {
        uint32_t w[HC / 2];
#pragma unroll
        for (int i = 0; i < HC / 2; ++i) {
          __nv_bfloat162 b2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
          w[i] = *reinterpret_cast<uint32_t*>(&b2);
        }
        if (col_ok) {
          uint4* d4 = reinterpret_cast<uint4*>(out_px + o * out_row_bytes);
          d4[0] = make_uint4(w[0], w[1], w[2], w[3]);
          if (HC == 16) d4[1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
        if (POOL) {                                // MaxPool2d(2): x pairs are adjacent lanes, y pairs consecutive rows
#pragma unroll
          for (int i = 0; i < HC / 2; ++i) {
            __nv_bfloat162 mx = *reinterpret_cast<__nv_bfloat162*>(&w[i]);
            const uint32_t ot = __shfl_xor_sync(0xffffffffu, w[i], 1);
            mx = __hmax2(mx, *reinterpret_cast<const __nv_bfloat162*>(&ot));
            if (o & 1) {
              mx = __hmax2(mx, *reinterpret_cast<__nv_bfloat162*>(&carry[i]));
              w[i] = *reinterpret_cast<uint32_t*>(&mx);
            } else {
              carry[i] = *reinterpret_cast<uint32_t*>(&mx);
            }
          }
          if ((o & 1) && col_ok && !(lane & 1)) {
            uint4* d4 = reinterpret_cast<uint4*>(pool_px + (o >> 1) * pool_row_bytes);
            d4[0] = make_uint4(w[0], w[1], w[2], w[3]);
            if (HC == 16) d4[1] = make_uint4(w[4], w[5], w[6], w[7]);
          }
        }
      } else {
        if (col_ok) {
          float4* d4 = reinterpret_cast<float4*>(out_px + o * out_row_bytes);
#pragma unroll
          for (int i = 0; i < HC / 4; ++i) d4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
        if (POOL) {
#pragma unroll
          for (int i = 0; i < HC; ++i) {
            float mx = fmaxf(v[i], __shfl_xor_sync(0xffffffffu, v[i], 1));
            if (o & 1) v[i] = fmaxf(mx, __uint_as_float(carry[i]));
            else carry[i] = __float_as_uint(mx);
          }
          if ((o & 1) && col_ok && !(lane & 1)) {
            float4* d4 = reinterpret_cast<float4*>(pool_px + (o >> 1) * pool_row_bytes);
#pragma unroll
            for (int i = 0; i < HC / 4; ++i) d4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
        }
      }
      s0 = s1; ph0 = ph1; s1 = s2; ph1 = ph2;
      if (++s2 == p.t_slots) { s2 = 0; ph2 ^= 1; }
    }
    // the last two input rows of the item were only read, never released: hand them back too
    __syncwarp();
    if (lane == 0) { mbar_arrive(&t_empty[s0]); mbar_arrive(&t_empty[s1]); }
    ts = s2; tph = ph2;                          // next item starts at the slot after its rows+2 input rows
  }
}

template <int ESZ, int KS>
__global__ void __launch_bounds__(kRowsThreads, 1) conv_rows_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                    const __grid_constant__ CUtensorMap tmW,
                                                                    const ConvRowsParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t a_full[kRowsMaxASlots], a_empty[kRowsMaxASlots];
  __shared__ uint64_t t_full[kRowsMaxTSlots], t_empty[kRowsMaxTSlots];
  __shared__ uint64_t w_full;
  __shared__ uint32_t tmem_slot;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  const uint32_t smem_off = ((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw);
  const uint32_t smem_base = smem_u32(smem_raw) + smem_off;
  const int w_tiles = p.kd * 3 * p.cin_chunks;
  const uint32_t a_base = smem_base + (uint32_t)w_tiles * p.w_tile_bytes;
  uint8_t* tail = smem_raw + smem_off + (size_t)w_tiles * p.w_tile_bytes + (size_t)p.a_slots * p.a_slot_bytes;
  float* s_scale = reinterpret_cast<float*>(tail);
  float* s_shift = s_scale + p.cp;
  float* s_headw = s_shift + p.cp;                        // [head_n][cp]
  float* s_part = s_headw + kMaxHead * p.cp;              // [2][128][kMaxHead]
  const uint32_t rb = p.row_bytes;
  const int nfold = 3 * p.cp;

  for (int i = threadIdx.x; i < p.cp; i += kRowsThreads) { s_scale[i] = p.scale[i]; s_shift[i] = p.shift[i]; }
  if (p.mode == EPI_HEAD)
    for (int i = threadIdx.x; i < p.head_n * p.cp; i += kRowsThreads) s_headw[i] = p.head_w[i];

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    for (int i = 0; i < p.a_slots; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < p.t_slots; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 8); }
    mbar_init(&w_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 2) {
    // ============================ folded weights: loaded once, resident ============================
    if (elect_one()) {
      mbar_arrive_expect_tx(&w_full, (uint32_t)w_tiles * (uint32_t)nfold * rb);
      for (int tdx = 0; tdx < p.kd * 3; ++tdx)                     // tdx = dz * 3 + dx
        for (int ch = 0; ch < p.cin_chunks; ++ch)
          asm volatile(
              "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
              "%5}], [%2];" ::"r"(smem_base + (uint32_t)(tdx * p.cin_chunks + ch) * p.w_tile_bytes),
              "l"(reinterpret_cast<uint64_t>(&tmW)), "r"(smem_u32(&w_full)), "r"(ch * p.ck), "r"(0), "r"(tdx)
              : "memory");
    }
  }
  if (warp == 0) {
    // ================================== input row (A) producer ===================================
    if (elect_one()) {
      const uint32_t row_tx = (uint32_t)p.cin_chunks * (uint32_t)kRowsPx * rb;
      int as = 0;
      uint32_t aph = 0;
      for (int t = blockIdx.x; t < p.total_items; t += gridDim.x) {
        const RowsItem it = rows_decode(p, t);
        for (int i = 0; i < it.rows + 2; ++i)
          for (int dz = 0; dz < p.kd; ++dz) {
            mbar_wait(&a_empty[as], aph ^ 1, 0xB00 + as);
            mbar_arrive_expect_tx(&a_full[as], row_tx);
            for (int ch = 0; ch < p.cin_chunks; ++ch)
              asm volatile(
                  "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
                  "%5, %6, %7}], [%2];" ::"r"(a_base + as * p.a_slot_bytes + ch * p.a_chunk_bytes),
                  "l"(reinterpret_cast<uint64_t>(&tmA)), "r"(smem_u32(&a_full[as])), "r"(ch * p.ck), "r"(it.x0 - 1),
                  "r"(it.y0 - 1 + i), "r"(it.z - (p.kd >> 1) + dz), "r"(it.b)
                  : "memory");
            if (++as == p.a_slots) { as = 0; aph ^= 1; }
          }
      }
    }
  } else if (warp == 1) {
    // ====================================== MMA issuer ======================================
    const uint32_t layout = rb == 128 ? 2u : (rb == 64 ? 4u : 6u);
    const uint32_t idesc = make_idesc(ESZ == 2 ? 1u : 2u, (uint32_t)nfold);
    const uint64_t a_desc0 = make_smem_desc(a_base, 8u * rb, layout);
    const uint64_t w_desc0 = make_smem_desc(smem_base, 8u * rb, layout);
    constexpr uint32_t px_step = 2u * KS;              // one pixel = row_bytes / 16
    const uint32_t aslot_step = p.a_slot_bytes >> 4, achunk_step = p.a_chunk_bytes >> 4, wtile_step = p.w_tile_bytes >> 4;
    mbar_wait(&w_full, 0, 0xC00);
    tc_fence_after();
    int as = 0, ts = 0;
    uint32_t aph = 0, tph = 0;
    for (int t = blockIdx.x; t < p.total_items; t += gridDim.x) {
      const RowsItem it = rows_decode(p, t);
      for (int i = 0; i < it.rows + 2; ++i) {
        mbar_wait(&t_empty[ts], tph ^ 1, 0xD00 + ts);              // epilogue released this TMEM slot
        tc_fence_after();
        const uint32_t tcol = tmem_base + (uint32_t)(ts * nfold);
        bool first = true;
        for (int dz = 0; dz < p.kd; ++dz) {
          mbar_wait(&a_full[as], aph, 0xE00 + as);
          tc_fence_after();
          for (int ch = 0; ch < p.cin_chunks; ++ch) {
            const uint64_t ad0 = a_desc0 + (uint64_t)(as * aslot_step + ch * achunk_step);
            const uint64_t wd0 = w_desc0 + (uint64_t)((dz * 3 * p.cin_chunks + ch) * wtile_step);
            const uint32_t wdx_step = p.cin_chunks * wtile_step;   // next dx tap
            if (elect_one()) {
              if (first) {
#pragma unroll
                for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                  for (int k = 0; k < KS; ++k) {
                    if (dx == 0 && k == 0) tc_mma_imm<ESZ, 0>(tcol, ad0 + (dx * px_step + 2 * k), wd0 + (dx * wdx_step + 2 * k), idesc);
                    else tc_mma_imm<ESZ, 1>(tcol, ad0 + (dx * px_step + 2 * k), wd0 + (dx * wdx_step + 2 * k), idesc);
                  }
              } else {
#pragma unroll
                for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                  for (int k = 0; k < KS; ++k)
                    tc_mma_imm<ESZ, 1>(tcol, ad0 + (dx * px_step + 2 * k), wd0 + (dx * wdx_step + 2 * k), idesc);
              }
            }
            first = false;
          }
          if (elect_one()) tc_commit(&a_empty[as]);
          if (++as == p.a_slots) { as = 0; aph ^= 1; }
        }
        if (elect_one()) tc_commit(&t_full[ts]);
        if (++ts == p.t_slots) { ts = 0; tph ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ======================================= epilogue =======================================
#define BIU_REPI(HC, MODE, POOL) \
    rows_epilogue<ESZ, HC, MODE, POOL>(p, tmem_base, t_full, t_empty, s_scale, s_shift, s_headw, s_part, warp, lane)
    if (p.cp == 32) {
      if (p.mode == EPI_HEAD) BIU_REPI(16, EPI_HEAD, false);
      else if (p.pool_out != nullptr) BIU_REPI(16, EPI_CONV, true);
      else BIU_REPI(16, EPI_CONV, false);
    } else {
      if (p.mode == EPI_HEAD) BIU_REPI(8, EPI_HEAD, false);
      else if (p.pool_out != nullptr) BIU_REPI(8, EPI_CONV, true);
      else BIU_REPI(8, EPI_CONV, false);
    }
#undef BIU_REPI
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

int rows_dispatch_bf16(const CUtensorMap& tmA, const CUtensorMap& tmW, const ConvRowsParams& p, int grid, int smem,
                       cudaStream_t stream);
int rows_dispatch_tf32(const CUtensorMap& tmA, const CUtensorMap& tmW, const ConvRowsParams& p, int grid, int smem,
                       cudaStream_t stream);

#define BIU_ROWS_LAUNCH(E, K)                                                                                     \
  do {                                                                                                             \
    static int max_set = 0;                                                                                        \
    if (smem > max_set) {                                                                                          \
      BIU_CHECK_CUDA(cudaFuncSetAttribute(conv_rows_kernel<E, K>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                          smem));                                                                  \
      max_set = smem;                                                                                              \
    }                                                                                                              \
    conv_rows_kernel<E, K><<<grid, kRowsThreads, smem, stream>>>(tmA, tmW, p);                                     \
  } while (0)
#define BIU_DEFINE_ROWS_DISPATCH(NAME, E)                                                                          \
  int NAME(const CUtensorMap& tmA, const CUtensorMap& tmW, const ConvRowsParams& p, int grid, int smem,            \
           cudaStream_t stream) {                                                                                  \
    const int ks = p.row_bytes / 32;                                                                               \
    if (ks == 4) BIU_ROWS_LAUNCH(E, 4); else if (ks == 2) BIU_ROWS_LAUNCH(E, 2); else BIU_ROWS_LAUNCH(E, 1);       \
    return 0;                                                                                                      \
  }

}  // namespace biu
