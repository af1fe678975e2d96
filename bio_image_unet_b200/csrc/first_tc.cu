// First block of every network (1 input channel: Conv 3x3(x3) + BatchNorm + LeakyReLU on the raw tile) on tcgen05.
//
// The CUDA-core form (direct.cu) spends 288 (2D, 32 filters) FMAs per pixel; here the 9 / 27 taps of a pixel are one
// K = 16 / 32 row of an im2col tile that four producer warps build directly in shared memory (no TMA: the source is
// a planar uint8 / float32 tile, 1 byte per pixel), and one MMA per 128 pixels does the arithmetic. uint8 inputs are
// exact in bf16 (integers 0..255): float32(u8) / 255 (unet/predict.py:192) becomes a factor 1/255 folded into the
// fp32 epilogue scale. Same tile walk (16 x 8*mt pixels), TMEM staging and epilogue (halo_epilogue: folded BN +
// LeakyReLU, transposed full-sector stores) as conv_halo.cuh. bf16 activations only; the tf32 / fp32 modes keep
// the CUDA-core kernel.
// STATUS (round 1): parity-green but measured slower than the CUDA-core kernel (cfg 2: 2.5 vs 1.7 ms per 200
// tiles; UNet3D(16): 2.9 vs 1.9 ms), so the network handle does not use it unless biu_net_set_first_tc(net, 1).
// Warp roles: 1 = TMEM alloc + MMA issuer, 4..11 = epilogue, 0 / 2 / 3 / 12 = im2col producers (128 threads = the
// 128 pixels of one MMA tile).
#include "conv_halo.cuh"
#include "launch.h"

namespace biu {

constexpr int kFirstTcThreads = 416;      // 13 warps

struct FirstTcExtra {
  const void* in;            // planar [B][1][D][H][W] uint8 or float32
  int in_kind;               // 0 = u8, 1 = f32
  const uint16_t* wgt;       // bf16 [n_blk][K], K = 16 (2D) or 32 (3D), tap-major, zero padded
  int stages;                // im2col ring depth
  uint32_t halo_off;         // byte offset (from the aligned smem base) of the producers' bf16 halo tiles [2][kd*18*(8mt+2)]
};

__device__ __forceinline__ void named_bar_sync_first(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int KS>           // K = 16 * KS: 1 (2D, 9 taps) or 2 (3D, 27 taps)
__global__ void __launch_bounds__(kFirstTcThreads, 1) first_tc_kernel(const ConvHaloParams p, const FirstTcExtra x) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t a_full[4], a_empty[4];
  __shared__ uint64_t acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;
  constexpr uint32_t RB = 32u * KS;                                  // bytes of one im2col row
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t smem_off = ((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw);
  const uint32_t smem_base = smem_u32(smem_raw) + smem_off;
  const uint32_t a_stage_bytes = (uint32_t)p.mt * 128u * RB;         // multiple of 4096
  const uint32_t b_base = smem_base + (uint32_t)x.stages * a_stage_bytes;
  const uint32_t b_bytes = (((uint32_t)p.n_blk * RB) + 1023u) & ~1023u;
  float* s_scale = reinterpret_cast<float*>(smem_raw + smem_off + x.stages * a_stage_bytes + b_bytes);
  float* s_shift = s_scale + p.n_total;
  float* s_headw = s_shift + p.n_total;
  const uint32_t stage_base = p.stage_bytes ? ((smem_u32(s_headw) + 15u) & ~15u) : 0u;
  const uint32_t acc_cols = (uint32_t)(p.mt * p.n_blk);
  uint32_t ncols = 32;
  while (ncols < 2 * acc_cols) ncols <<= 1;

  for (int i = threadIdx.x; i < p.n_total; i += kFirstTcThreads) { s_scale[i] = p.scale[i]; s_shift[i] = p.shift[i]; }
  // weights -> shared memory in the swizzled K-major layout the MMA descriptor expects (row n: RB bytes, 16-byte
  // chunk c stored at c ^ f(n): 32B swizzle f = (n >> 2) & 1, 64B swizzle f = (n >> 1) & 3)
  for (int i = threadIdx.x; i < p.n_blk * (int)(RB / 16); i += kFirstTcThreads) {
    const int nrow = i / (int)(RB / 16), c = i % (int)(RB / 16);
    const uint32_t sw = KS == 1 ? ((uint32_t)(nrow >> 2) & 1u) : ((uint32_t)(nrow >> 1) & 3u);
    const uint4 v = *reinterpret_cast<const uint4*>(x.wgt + (size_t)nrow * (RB / 2) + c * 8);
    *reinterpret_cast<uint4*>(smem_raw + smem_off + x.stages * a_stage_bytes + nrow * RB + (((uint32_t)c ^ sw) << 4)) = v;
  }
  if (warp == 0 && elect_one()) {
    for (int i = 0; i < x.stages; ++i) { mbar_init(&a_full[i], 4); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_slot, ncols);
    tmem_relinquish();
  }
  fence_proxy_async();                 // the generic-proxy weight stores above become visible to the tensor core
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 1) {
    // ====================================== MMA issuer ======================================
    const uint32_t layout = KS == 1 ? 6u : 4u;                       // SWIZZLE_32B / SWIZZLE_64B
    const uint32_t idesc = make_idesc(1u, (uint32_t)p.n_blk);
    const uint64_t a_desc0 = make_smem_desc(smem_base, 8u * RB, layout);
    const uint64_t b_desc0 = make_smem_desc(b_base, 8u * RB, layout);
    int as = 0, it = 0;
    uint32_t aph = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
      const int acs = it & 1;
      mbar_wait(&acc_empty[acs], ((it >> 1) & 1) ^ 1, 0x900 + acs);
      mbar_wait(&a_full[as], aph, 0x600 + as);
      tc_fence_after();
      const uint32_t tacc = tmem_base + acs * acc_cols;
      const uint64_t ad0 = a_desc0 + (uint64_t)((as * a_stage_bytes) >> 4);
      if (elect_one()) {
        for (int j = 0; j < p.mt; ++j) {
#pragma unroll
          for (int k = 0; k < KS; ++k) {
            const uint64_t ad = ad0 + (uint64_t)(((uint32_t)j * 128u * RB) >> 4) + 2 * k;
            if (k == 0) tc_mma_imm<2, 0>(tacc + j * p.n_blk, ad, b_desc0 + 2 * k, idesc);
            else tc_mma_imm<2, 1>(tacc + j * p.n_blk, ad, b_desc0 + 2 * k, idesc);
          }
        }
        tc_commit(&a_empty[as]);
        tc_commit(&acc_full[acs]);
      }
      if (++as == x.stages) { as = 0; aph ^= 1; }
    }
  } else if (warp >= 4 && warp < 12) {
    // ======================================= epilogue =======================================
    if (p.n_blk % 32 == 0)
      halo_epilogue<2, 32, EPI_CONV, false>(p, tmem_base, acc_cols, acc_full, acc_empty, s_scale, s_shift, s_headw,
                                            stage_base, warp, lane);
    else
      halo_epilogue<2, 16, EPI_CONV, false>(p, tmem_base, acc_cols, acc_full, acc_empty, s_scale, s_shift, s_headw,
                                            stage_base, warp, lane);
  } else {
    // ================================= im2col producers (warps 0, 2, 3, 12) ==================================
    // Per macro tile: (1) the 128 threads fetch the (kd x 18 x (8*mt+2)) halo of the tile from global memory,
    // already one tile ahead, as independent predicated loads (zero outside the image = the conv padding) and park
    // it in shared memory as bf16; (2) every thread builds the im2col rows of its pixel of each MMA tile from there.
    const int pidx = warp == 0 ? 0 : (warp == 12 ? 3 : warp - 1);
    const int tid = pidx * 32 + lane;                                // 0..127; also the row m of the MMA tile
    const int m = tid;
    constexpr int P = KS == 1 ? 1 : 3;                               // planes of the halo
    constexpr int R = 18;
    const int cw = 8 * p.mt + 2;
    const int elems = P * R * cw;
    constexpr int kMaxPer = KS == 1 ? 10 : 28;                       // ceil(P * 18 * 66 / 128)
    uint16_t* s_halo = reinterpret_cast<uint16_t*>(smem_raw + smem_off + x.halo_off);     // [2][elems]
    const long long plane = (long long)p.H * p.W;
    const uint8_t* in8 = reinterpret_cast<const uint8_t*>(x.in);
    const float* inf = reinterpret_cast<const float*>(x.in);
    uint16_t pre[kMaxPer];
    // tile-independent part of the halo element -> (plane, row, column) split, once per thread: packed as
    // column | row << 8 | plane << 16 (0xffffffff: no such element)
    uint32_t rel[kMaxPer];
#pragma unroll
    for (int i = 0; i < kMaxPer; ++i) {
      const int e = tid + i * 128;
      const int cx = e % cw, rr = e / cw;
      rel[i] = e < elems ? ((uint32_t)cx | ((uint32_t)(rr % R) << 8) | ((uint32_t)(rr / R) << 16)) : 0xffffffffu;
    }
    auto prefetch = [&](int t) {
      const HaloTile tl = halo_decode(p, t);
      const long long img = (long long)tl.b0 * p.D * plane;
#pragma unroll
      for (int i = 0; i < kMaxPer; ++i) {
        float v = 0.f;
        if (rel[i] != 0xffffffffu) {
          const int xx = tl.x0 - 1 + (int)(rel[i] & 0xff), yy = tl.y0 - 1 + (int)((rel[i] >> 8) & 0xff);
          const int zz = tl.z0 - (P >> 1) + (int)(rel[i] >> 16);
          if ((unsigned)xx < (unsigned)p.W && (unsigned)yy < (unsigned)p.H && (unsigned)zz < (unsigned)p.D) {
            const long long off = img + (long long)zz * plane + yy * p.W + xx;
            v = x.in_kind == 0 ? (float)__ldg(in8 + off) : __ldg(inf + off);
          }
        }
        pre[i] = __bfloat16_as_ushort(__float2bfloat16_rn(v));
      }
    };
    const int hrow = (m >> 3) * cw + (m & 7);                        // this pixel's top-left tap in the halo tile
    int as = 0, buf = 0;
    uint32_t aph = 0;
    if ((int)blockIdx.x < p.total_tiles) prefetch(blockIdx.x);
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      uint16_t* hb = s_halo + buf * elems;
#pragma unroll
      for (int i = 0; i < kMaxPer; ++i) {
        const int e = tid + i * 128;
        if (e < elems) hb[e] = pre[i];
      }
      named_bar_sync_first(1, 128);
      if (t + (int)gridDim.x < p.total_tiles) prefetch(t + gridDim.x);          // next tile's loads fly during the im2col
      mbar_wait(&a_empty[as], aph ^ 1, 0x400 + as);
      uint8_t* stage = smem_raw + smem_off + as * a_stage_bytes;
      const uint32_t sw = KS == 1 ? ((uint32_t)(m >> 2) & 1u) : ((uint32_t)(m >> 1) & 3u);
      for (int j = 0; j < p.mt; ++j) {
        uint32_t pk[8 * KS];
#pragma unroll
        for (int i = 0; i < 8 * KS; ++i) pk[i] = 0u;
#pragma unroll
        for (int dz = 0; dz < P; ++dz)
#pragma unroll
          for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              const int tap = (dz * 3 + dy) * 3 + dx;
              const uint32_t h = hb[hrow + (dz * R + dy) * cw + 8 * j + dx];
              pk[tap >> 1] |= (tap & 1) ? (h << 16) : h;
            }
        uint8_t* row = stage + (size_t)j * 128 * RB + (size_t)m * RB;
#pragma unroll
        for (int c = 0; c < 2 * KS; ++c)
          *reinterpret_cast<uint4*>(row + (((uint32_t)c ^ sw) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
      }
      fence_proxy_async();             // generic-proxy smem writes -> visible to the async proxy (tcgen05.mma reads)
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[as]);
      if (++as == x.stages) { as = 0; aph ^= 1; }
      buf ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, ncols);
}

// scale / 255 for uint8 tiles (float32(u8) / 255, unet/predict.py:192), cached per (scale pointer)
__global__ void scale_div255_kernel(const float* __restrict__ s, float* __restrict__ o, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) o[i] = __fdiv_rn(s[i], 255.0f);
}

bool first_tc_supported(const FirstConvArgs& a) {
  return a.esz == 2 && a.cin == 1 && a.wgt_tc != nullptr && a.scale255 != nullptr && (a.kd == 1 || a.kd == 3) &&
         a.cout_pad % 16 == 0 && a.cout_pad <= 64 && a.H >= 16 && a.W >= 8 && a.out_ctot % 8 == 0 && a.out_coff % 8 == 0;
}

int launch_first_tc(const FirstConvArgs& a, cudaStream_t stream) {
  BIU_REQUIRE(first_tc_supported(a), "first_tc: unsupported configuration");
  ConvHaloParams p;
  memset(&p, 0, sizeof(p));
  p.W = a.W; p.H = a.H; p.D = a.D; p.B = a.B;
  int mt = 256 / a.cout_pad;
  if (mt > 8) mt = 8;
  const int w8 = (a.W + 7) / 8;
  if (mt > w8) mt = w8;
  while (mt & (mt - 1)) --mt;
  p.mt = mt;
  p.tiles_x = ceil_div(a.W, 8 * mt); p.tiles_y = ceil_div(a.H, 16);
  p.n_blocks = 1; p.total_tiles = p.tiles_x * p.tiles_y * a.D * a.B;
  p.kd = a.kd; p.halo = 1;
  p.n_blk = a.cout_pad; p.n_total = a.cout_pad;
  p.mode = EPI_CONV; p.slope = a.slope;
  p.scale = a.in_kind == 0 ? a.scale255 : a.scale; p.shift = a.shift;
  p.out = a.out; p.out_ctot = a.out_ctot; p.out_coff = a.out_coff;
  p.stage_bytes = (a.cout_pad % 32 == 0) ? 8 * 2048 : 0;
  FirstTcExtra x;
  x.in = a.in; x.in_kind = a.in_kind; x.wgt = reinterpret_cast<const uint16_t*>(a.wgt_tc);
  const int ks = a.kd == 1 ? 1 : 2;
  const int a_stage = mt * 128 * 32 * ks;
  x.stages = 3;
  const int b_bytes = ((a.cout_pad * 32 * ks) + 1023) & ~1023;
  const int halo_bytes = 2 * (a.kd * 18 * (8 * mt + 2)) * 2;
  x.halo_off = (uint32_t)((x.stages * a_stage + b_bytes + 2 * a.cout_pad * 4 + 64 + p.stage_bytes + 15) & ~15);
  const int smem = (int)x.halo_off + halo_bytes + 1024;
  static int dev_sms = 0;
  if (dev_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, dev);
    if (dev_sms <= 0) dev_sms = 148;
  }
  const int grid = p.total_tiles < dev_sms ? p.total_tiles : dev_sms;
  if (ks == 1) {
    static int set1 = 0;
    if (smem > set1) { BIU_CHECK_CUDA(cudaFuncSetAttribute(first_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); set1 = smem; }
    first_tc_kernel<1><<<grid, kFirstTcThreads, smem, stream>>>(p, x);
  } else {
    static int set2 = 0;
    if (smem > set2) { BIU_CHECK_CUDA(cudaFuncSetAttribute(first_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); set2 = smem; }
    first_tc_kernel<2><<<grid, kFirstTcThreads, smem, stream>>>(p, x);
  }
  BIU_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int launch_scale_div255(const float* s, float* o, int n, cudaStream_t stream) {
  scale_div255_kernel<<<(n + 127) / 128, 128, 0, stream>>>(s, o, n);
  BIU_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace biu
