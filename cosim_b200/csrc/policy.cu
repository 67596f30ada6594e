// policy.cu -- the batched policy MLP on 5th-gen tensor cores (tcgen05 + TMEM), sm_100a only.
//
// Replaces MLPPolicy.get_action (/root/reference/core/policy.py:11-21): action = clip(MLP(state), -1, 1),
// where the reference runs an ONNX graph through onnxruntime's CPU provider one state at a time.
//
// One CTA (256 threads) owns a tile of 128 environments = the 128 TMEM lanes:
//   * the activation tile [128 x K] lives in shared memory as bf16 in the UMMA "no-swizzle, K-major"
//     canonical layout (8x8 core matrices of 128 B), and is overwritten in place layer after layer --
//     the accumulators of a whole layer (<= 512 fp32 columns) sit in TMEM, so the input of a layer is
//     dead by the time its output is written back;
//   * weights are pre-packed on the host into [n-chunk][k-chunk] blocks (<= 128 x 64 bf16 = 16 KB) already
//     in canonical order, so a block is ONE contiguous TMA bulk copy (cp.async.bulk -> UBLKCP) into a
//     2-stage ring, tracked by mbarriers;
//   * thread 0 issues tcgen05.mma (M = 128, N = chunk width, K = 16 per instruction), tcgen05.commit
//     releases the ring slot / publishes the accumulators;
//   * all 8 warps run the epilogue: tcgen05.ld (32 lanes x 16 columns), + bias, ELU/tanh/ReLU, bf16
//     pack, store to the activation tile; the last layer clips to [-1, 1] and writes fp32 actions.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include <string>
#include <vector>
#include "../../include/cosim_b200.h"

#define POL_MAX_LAYERS 6
#define POL_MAX_WIDTH 512
#define POL_TILE_M 128
#define POL_KC 64                       // k-chunk (elements)
#define POL_NC 128                      // n-chunk (max rows of a weight block)
#define POL_WSTAGE_BYTES (POL_NC * POL_KC * 2)
#define POL_ACT_BYTES (POL_TILE_M * POL_MAX_WIDTH * 2)
#define POL_SMEM_BYTES (POL_ACT_BYTES + 2 * POL_WSTAGE_BYTES + 64)

struct PolicyDev {
  int nlayers, act, in_dim, out_dim, raw_out;      // raw_out: 0 = clip(-1, 1), 1 = plain linear output, 2 = hidden activation applied
  int K[POL_MAX_LAYERS], N[POL_MAX_LAYERS];        // padded: K % 64 == 0, N % 64 == 0 (hidden) or % 16 == 0 (last)
  const __nv_bfloat16* w[POL_MAX_LAYERS];          // packed blocks
  const float* b[POL_MAX_LAYERS];                  // padded biases
};

// ------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t spin = 0; !ok; ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (spin > (1u << 24)) { printf("cosim policy: mbarrier wait timed out (bar %u parity %u)\n", bar, parity); __trap(); }
  }
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t* r) {
  // load + wait in ONE asm statement so that no use of r[] can be scheduled between them
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
               "tcgen05.wait::ld.sync.aligned;"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
}
// UMMA shared-memory descriptor, SWIZZLE_NONE, K-major (cute::UMMA::SmemDescriptor): start address, leading
// byte offset (between core matrices along K), stride byte offset (between 8-row groups), version = 1
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = BF16, both K-major, M = 128
__device__ __forceinline__ uint32_t umma_idesc(int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(POL_TILE_M >> 4) << 24); }

__device__ __forceinline__ float activate(float x, int act) {
  if (act == 0) return x > 0.f ? x : __expf(x) - 1.f;   // ELU(alpha = 1); the result is rounded to bf16 (2^-8) right after
  if (act == 1) return tanhf(x);
  return fmaxf(x, 0.f);                               // ReLU
}

// ------------------------------------------------------------------------------------------ kernel
__global__ void __launch_bounds__(256, 1) k_policy_mlp(const __grid_constant__ PolicyDev p, const float* __restrict__ state, int num_envs, float* __restrict__ action) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* act_buf = smem;                                   // [k/8][row/8][row%8][k%8] bf16
  uint8_t* wbuf = smem + POL_ACT_BYTES;                      // 2 stages
  uint64_t* bars = (uint64_t*)(smem + POL_ACT_BYTES + 2 * POL_WSTAGE_BYTES);   // full[2], empty[2], acc
  __shared__ uint32_t tmem_base_sh;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + 2), bar_acc = smem_u32(bars + 4);

  if (tid == 0) {
    mbar_init(bar_full, 1); mbar_init(bar_full + 8, 1); mbar_init(bar_empty, 1); mbar_init(bar_empty + 8, 1); mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_sh;

  uint32_t g = 0;        // weight blocks streamed so far (thread 0)
  uint32_t nacc = 0;     // layers finished so far (all threads)
  const int ntiles = (num_envs + POL_TILE_M - 1) / POL_TILE_M;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    // ---- stage the input tile: fp32 state rows -> bf16 canonical layout (one warp per row, lanes over 8-column groups)
    const int K0 = p.K[0];
    for (int r = warp; r < POL_TILE_M; r += 8) {
      const int env = tile * POL_TILE_M + r;
      const float* src = state + (size_t)env * p.in_dim;
      for (int j = tid & 31; j < K0 / 8; j += 32) {
        __nv_bfloat162 v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c = j * 8 + 2 * q;
          const float a = (env < num_envs && c < p.in_dim) ? __ldg(src + c) : 0.f;
          const float b = (env < num_envs && c + 1 < p.in_dim) ? __ldg(src + c + 1) : 0.f;
          v[q] = __floats2bfloat162_rn(a, b);
        }
        *(uint4*)(act_buf + ((size_t)(j * 16 + (r >> 3)) * 128 + (r & 7) * 16)) = *(uint4*)v;
      }
    }
    proxy_fence_async();
    __syncthreads();

    for (int l = 0; l < p.nlayers; ++l) {
      const int K = p.K[l], Nl = p.N[l];
      const bool last = (l == p.nlayers - 1);
      const int nkc = K / POL_KC;
      // hidden layers are at most 512 wide (one TMEM pass).  The LAST layer writes to global memory, so its input tile
      // stays intact and wider outputs (LSTM gate pre-activations: 4 H columns) run as several 512-column passes.
      for (int nbase = 0; nbase < Nl; nbase += POL_MAX_WIDTH) {
      const int N = min(POL_MAX_WIDTH, Nl - nbase);
      const int nnc = (N + POL_NC - 1) / POL_NC;
      if (tid == 0) {
        tc_fence_after();
        const int C = nkc * nnc;
        auto issue_load = [&](int c) {
          const uint32_t gg = g + c, s = gg & 1, u = gg >> 1;
          if (u >= 1) mbar_wait(bar_empty + 8 * s, (u - 1) & 1);
          const int nc = c / nkc, kc = c - nc * nkc;
          const int n0 = nbase + nc * POL_NC, Nc = min(POL_NC, Nl - n0);
          const uint32_t bytes = (uint32_t)Nc * POL_KC * 2;
          const __nv_bfloat16* src = p.w[l] + ((size_t)n0 * K + (size_t)kc * Nc * POL_KC);
          mbar_expect_tx(bar_full + 8 * s, bytes);
          tma_bulk_g2s(smem_u32(wbuf + s * POL_WSTAGE_BYTES), src, bytes, bar_full + 8 * s);
        };
        issue_load(0);
        for (int c = 0; c < C; ++c) {
          if (c + 1 < C) issue_load(c + 1);
          const uint32_t gg = g + c, s = gg & 1, u = gg >> 1;
          mbar_wait(bar_full + 8 * s, u & 1);
          tc_fence_after();
          const int nc = c / nkc, kc = c - nc * nkc;
          const int n0 = nbase + nc * POL_NC, Nc = min(POL_NC, Nl - n0);
          const uint32_t idesc = umma_idesc(Nc);
          const uint32_t a_base = smem_u32(act_buf) + (uint32_t)(kc * (POL_KC / 8)) * 16 * 128;
          const uint32_t b_base = smem_u32(wbuf + s * POL_WSTAGE_BYTES);
#pragma unroll
          for (int kk = 0; kk < POL_KC / 16; ++kk) {
            const uint64_t ad = umma_desc(a_base + (uint32_t)kk * 2 * 16 * 128, 16 * 128, 128);
            const uint64_t bd = umma_desc(b_base + (uint32_t)kk * 2 * (Nc / 8) * 128, (uint32_t)(Nc / 8) * 128, 128);
            tc_mma_bf16(tmem_base + (uint32_t)(n0 - nbase), ad, bd, idesc, (kc > 0 || kk > 0) ? 1u : 0u);
          }
          tc_commit(bar_empty + 8 * s);
        }
        tc_commit(bar_acc);
        g += C;
      }
      // ---- epilogue (all threads)
      mbar_wait(bar_acc, nacc & 1);
      ++nacc;
      tc_fence_after();
      // 8 warps: warp w reads TMEM lanes 32 (w % 4) .. +31 (a warp can only reach its own lane quarter); the two warps
      // of a quarter alternate 16-column chunks
      const int row = (warp & 3) * 32 + (tid & 31), env = tile * POL_TILE_M + row;
      const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
      for (int c0 = (warp >> 2) * 16; c0 < N; c0 += 32) {
        uint32_t r[16];
        tc_ld16(lane_addr + (uint32_t)c0, r);
        if (!last) {
          __nv_bfloat162 v[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float a = activate(__uint_as_float(r[2 * q]) + __ldg(p.b[l] + c0 + 2 * q), p.act);
            const float b = activate(__uint_as_float(r[2 * q + 1]) + __ldg(p.b[l] + c0 + 2 * q + 1), p.act);
            v[q] = __floats2bfloat162_rn(a, b);
          }
          const int j = c0 >> 3;
          *(uint4*)(act_buf + ((size_t)(j * 16 + (row >> 3)) * 128 + (row & 7) * 16)) = *(uint4*)&v[0];
          *(uint4*)(act_buf + ((size_t)((j + 1) * 16 + (row >> 3)) * 128 + (row & 7) * 16)) = *(uint4*)&v[4];
        } else if (env < num_envs) {
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const int c = nbase + c0 + q;
            if (c < p.out_dim) {
              const float y = __uint_as_float(r[q]) + __ldg(p.b[l] + c);
              action[(size_t)env * p.out_dim + c] = p.raw_out == 0 ? fminf(1.f, fmaxf(-1.f, y)) : (p.raw_out == 2 ? activate(y, p.act) : y);
            }
          }
        }
      }
      tc_fence_before();
      proxy_fence_async();
      __syncthreads();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}

// LSTM cell update (ONNX LSTM op, gate order i, o, f, c; used by LSTMPolicy, core/policy.py:24-47): gates = W x + R h + b
// come from k_policy_mlp with raw output; c and h are updated in place.
__global__ void k_lstm_cell(const float* __restrict__ gates, float* __restrict__ c, float* __restrict__ h, int num_envs, int H) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)num_envs * H) return;
  const int e = (int)(t / H), j = (int)(t - (long long)e * H);
  const float* g = gates + (size_t)e * 4 * H;
  const float gi = 1.f / (1.f + __expf(-g[j])), go = 1.f / (1.f + __expf(-g[H + j])), gf = 1.f / (1.f + __expf(-g[2 * H + j])), gc = tanhf(g[3 * H + j]);
  const float cn = gf * c[t] + gi * gc;
  c[t] = cn;
  h[t] = go * tanhf(cn);
}

// ------------------------------------------------------------------------------------------ host side
struct cosim_policy {
  PolicyDev d;
  int device = 0, launches = 0, sms = 148;
  std::vector<void*> allocs;
};

static uint16_t f2bf(float f) {   // round-to-nearest-even, as __float2bfloat16_rn
  uint32_t x; memcpy(&x, &f, 4);
  if ((x & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((x >> 16) | 0x40);
  x += 0x7fffu + ((x >> 16) & 1u);
  return (uint16_t)(x >> 16);
}

extern "C" {

int cosim_policy_create(int device, int nlayers, const int* dims, const float* const* weights_host, const float* const* biases_host,
                        int activation, cosim_policy** out) {
  const int raw_out = (activation >> 8) & 3;       // 1 = COSIM_POLICY_RAW_OUTPUT, 2 = COSIM_POLICY_ACTIVATED_OUTPUT
  activation &= 0xff;
  if (!out || !dims || !weights_host || !biases_host || nlayers < 1 || nlayers > POL_MAX_LAYERS || activation < 0 || activation > 2) return COSIM_ERR_ARG;
  *out = nullptr;
  for (int l = 0; l <= nlayers; ++l) {
    const int lim = (l == nlayers) ? 4 * POL_MAX_WIDTH : POL_MAX_WIDTH;       // only the output may be wider than one TMEM pass
    if (dims[l] < 1 || dims[l] > lim) { fprintf(stderr, "cosim_policy_create: layer width %d outside [1, %d]\n", dims[l], lim); return COSIM_ERR_ARG; }
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { fprintf(stderr, "cosim_b200: no CUDA device -- the policy has no CPU path\n"); return COSIM_ERR_CUDA; }
  if (cudaSetDevice(device) != cudaSuccess) return COSIM_ERR_CUDA;
  cosim_policy* p = new cosim_policy;
  p->device = device;
  memset(&p->d, 0, sizeof(p->d));
  p->d.nlayers = nlayers; p->d.act = activation; p->d.in_dim = dims[0]; p->d.out_dim = dims[nlayers]; p->d.raw_out = raw_out;
  cudaDeviceProp prop; if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) p->sms = prop.multiProcessorCount;
  for (int l = 0; l < nlayers; ++l) {
    const int kin = dims[l], nout = dims[l + 1];
    const int K = (kin + 63) / 64 * 64, N = (l == nlayers - 1) ? (nout + 15) / 16 * 16 : (nout + 63) / 64 * 64;
    p->d.K[l] = K; p->d.N[l] = N;
    std::vector<uint16_t> packed((size_t)N * K, 0);
    std::vector<float> bias(N, 0.f);
    const int nkc = K / POL_KC;
    for (int n0 = 0; n0 < N; n0 += POL_NC) {
      const int Nc = (N - n0) < POL_NC ? (N - n0) : POL_NC;
      for (int kc = 0; kc < nkc; ++kc) {
        uint16_t* blk = packed.data() + (size_t)n0 * K + (size_t)kc * Nc * POL_KC;
        for (int n = 0; n < Nc; ++n) for (int k = 0; k < POL_KC; ++k) {
          const int gn = n0 + n, gk = kc * POL_KC + k;
          const float w = (gn < nout && gk < kin) ? weights_host[l][(size_t)gn * kin + gk] : 0.f;
          blk[((size_t)(k / 8) * (Nc / 8) + n / 8) * 64 + (n % 8) * 8 + (k % 8)] = f2bf(w);
        }
      }
    }
    for (int n = 0; n < nout; ++n) bias[n] = biases_host[l][n];
    void *dw = nullptr, *db = nullptr;
    if (cudaMalloc(&dw, packed.size() * 2) != cudaSuccess || cudaMalloc(&db, bias.size() * 4) != cudaSuccess) { for (void* q : p->allocs) cudaFree(q); delete p; return COSIM_ERR_CUDA; }
    p->allocs.push_back(dw); p->allocs.push_back(db);
    cudaMemcpy(dw, packed.data(), packed.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice);
    p->d.w[l] = (const __nv_bfloat16*)dw; p->d.b[l] = (const float*)db;
  }
  if (cudaFuncSetAttribute(k_policy_mlp, cudaFuncAttributeMaxDynamicSharedMemorySize, POL_SMEM_BYTES) != cudaSuccess) {
    fprintf(stderr, "cosim_policy_create: cannot reserve %d B of shared memory: %s\n", POL_SMEM_BYTES, cudaGetErrorString(cudaGetLastError()));
    for (void* q : p->allocs) cudaFree(q); delete p; return COSIM_ERR_CUDA;
  }
  *out = p;
  return COSIM_OK;
}

void cosim_policy_destroy(cosim_policy* p) {
  if (!p) return;
  cudaSetDevice(p->device);
  for (void* q : p->allocs) cudaFree(q);
  delete p;
}

int cosim_policy_forward(cosim_policy* p, const float* state, int num_envs, float* action_out, void* stream) {
  if (!p || !state || !action_out || num_envs <= 0) return COSIM_ERR_ARG;
  const int ntiles = (num_envs + POL_TILE_M - 1) / POL_TILE_M;
  const int grid = ntiles < p->sms ? ntiles : p->sms;
  k_policy_mlp<<<grid, 256, POL_SMEM_BYTES, (cudaStream_t)stream>>>(p->d, state, num_envs, action_out);
  p->launches++;
  return cudaGetLastError() == cudaSuccess ? COSIM_OK : COSIM_ERR_CUDA;
}
int cosim_policy_launch_count(const cosim_policy* p) { return p ? p->launches : COSIM_ERR_ARG; }

int cosim_lstm_cell(const float* gates, float* c, float* h, int num_envs, int hidden, void* stream) {
  if (!gates || !c || !h || num_envs <= 0 || hidden <= 0) return COSIM_ERR_ARG;
  const long long n = (long long)num_envs * hidden;
  k_lstm_cell<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(gates, c, h, num_envs, hidden);
  return cudaGetLastError() == cudaSuccess ? COSIM_OK : COSIM_ERR_CUDA;
}

}  // extern "C"
