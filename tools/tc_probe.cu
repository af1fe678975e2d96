// Probe: which shared-memory descriptor forms does tcgen05.mma accept for a 128B-swizzled K-major A operand
// whose start address is NOT 1024-byte aligned (row-shifted views of one TMA-loaded halo tile)?
// Prints max |err| per case; cases that hang are bounded by the mbarrier timeout in common.cuh.
#include "../bio_image_unet_b200/csrc/common.cuh"
#include <cudaTypedefs.h>
#include <vector>
#include <cstdlib>
#include <cmath>
using namespace biu;
namespace biu { void set_error(const char*, ...) {} const char* get_error() { return ""; } }

constexpr int ROWS = 192;   // rows of A staged in smem
constexpr int N = 32;

struct Case { int start_rows; int sbo_bytes; int base_off_mode; int rb; };  // base_off_mode: 0 -> 0, 1 -> (addr>>7)&7

__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap tmA,
                                                    const __grid_constant__ CUtensorMap tmB, Case c, float* out) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t bar_full, bar_done;
  __shared__ uint32_t tslot;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t a_s = base, b_s = base + ROWS * 128;
  const uint32_t RB = c.rb;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && elect_one()) { mbar_init(&bar_full, 1); mbar_init(&bar_done, 1); fence_mbar_init(); }
  if (warp == 1) { tmem_alloc(&tslot, 32); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tslot;
  if (warp == 0 && elect_one()) {
    mbar_arrive_expect_tx(&bar_full, ROWS * RB + N * RB);
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(a_s), "l"(reinterpret_cast<uint64_t>(&tmA)), "r"(smem_u32(&bar_full)), "r"(0), "r"(0) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(b_s), "l"(reinterpret_cast<uint64_t>(&tmB)), "r"(smem_u32(&bar_full)), "r"(0), "r"(0) : "memory");
    mbar_wait(&bar_full, 0, 1);
    tc_fence_after();
    const uint32_t idesc = make_idesc(1, N);
    const uint32_t a_addr = a_s + c.start_rows * RB;
    const uint32_t bo = c.base_off_mode ? ((a_addr >> 7) & 7) : 0;
    const uint32_t lay = RB == 128 ? 2 : (RB == 64 ? 4 : 6);
    for (int k = 0; k < (int)(RB / 32); ++k) {
      uint64_t ad = make_smem_desc(a_addr + k * 32, c.sbo_bytes, lay, bo);
      uint64_t bd = make_smem_desc(b_s + k * 32, 8 * RB, lay, 0);
      tc_mma_f16(tm, ad, bd, idesc, k != 0);
    }
    tc_commit(&bar_done);
  }
  __syncwarp();
  mbar_wait(&bar_done, 0, 2);
  tc_fence_after();
  uint32_t r[16];
  for (int c0 = 0; c0 < N; c0 += 16) {
    tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * N + c0 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) tmem_dealloc(tm, 32);
}

static uint16_t f2bf(float v) { uint32_t u; memcpy(&u, &v, 4); u += 0x7FFF + ((u >> 16) & 1); return u >> 16; }
static float bf2f(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }

int main() {
  for (int RB : {128, 64, 32}) {
  const int K = RB / 2;
  printf("==== row bytes %d (swizzle %dB), K = %d\n", RB, RB, K);
  std::vector<uint16_t> A(ROWS * K), B(N * K);
  srand(1);
  for (auto& v : A) v = f2bf((rand() % 201 - 100) / 64.0f);
  for (auto& v : B) v = f2bf((rand() % 201 - 100) / 64.0f);
  uint16_t *dA, *dB; float* dO;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dO, 128 * N * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  void* sym = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
  auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(sym);
  CUtensorMap tmA, tmB;
  cuuint64_t dimsA[2] = {(cuuint64_t)K, ROWS}, dimsB[2] = {(cuuint64_t)K, N}, strides[1] = {(cuuint64_t)RB};
  cuuint32_t boxA[2] = {(cuuint32_t)K, ROWS}, boxB[2] = {(cuuint32_t)K, N}, es[2] = {1, 1};
  CUtensorMapSwizzle sw = RB == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (RB == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  int r1 = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dimsA, strides, boxA, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  int r2 = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dimsB, strides, boxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode %d %d\n", r1, r2);
  const int smem = ROWS * 128 + N * 128 + 2048;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> O(128 * N);
  // (a) contiguous rows, shifted start: row m -> smem row start+m
  // (b) halo walk: 16 rows x 8 cols out of a pitch-10 tile: row m=(r,c) -> smem row (r+dy)*10 + c+dx ; SBO = 1280
  struct Spec { int start, sbo, mode, kind, dy, dx; };
  std::vector<Spec> specs;
  for (int s : {0, 8, 1, 2, 3, 7, 9, 17}) for (int mode : {0}) specs.push_back({s, 8 * RB, mode, 0, 0, 0});
  for (int dy = 0; dy < 3; ++dy) for (int dx = 0; dx < 3; ++dx) for (int mode : {0})
    specs.push_back({dy * 10 + dx, 10 * RB, mode, 1, dy, dx});
  for (auto& sp : specs) {
    Case c{sp.start, sp.sbo, sp.mode, RB};
    cudaMemset(dO, 0, 128 * N * 4);
    probe_kernel<<<1, 128, smem>>>(tmA, tmB, c, dO);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("case start=%d sbo=%d mode=%d kind=%d: CUDA error %s\n", sp.start, sp.sbo, sp.mode, sp.kind, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (int m = 0; m < 128; ++m) {
      int row = sp.kind == 0 ? sp.start + m : ((m / 8) + sp.dy) * 10 + (m % 8) + sp.dx;
      for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int k = 0; k < K; ++k) ref += (double)bf2f(A[row * K + k]) * bf2f(B[n * K + k]);
        maxerr = fmax(maxerr, fabs(ref - O[m * N + n]));
      }
    }
    printf("kind=%s start_rows=%2d sbo=%4d base_offset=%s : max_err=%.5f %s\n", sp.kind ? "halo" : "shift", sp.start, sp.sbo,
           sp.mode ? "(addr>>7)&7" : "0", maxerr, maxerr < 1e-3 ? "OK" : "WRONG");
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dO);
  }
  return 0;
}
