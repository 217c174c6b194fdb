// hk_actor.cuh -- the reference's TD3 actor (ActorNetwork.forward, rl/td3/networks.py:17-20:
// tanh(W3 tanh(W2 tanh(W1 obs + b1) + b2) + b3), 18 -> 256 -> 256 -> 4) as ONE fused sm_100a kernel on the 5th-generation
// tensor cores (BASELINE config 5: the actor consumes the env's observation tensor in place and writes the action tensor
// the next hk_step reads).
//
// This is the one GEMM-shaped op next to the env path, so it runs on tcgen05: a CTA owns 128 observation rows at a time
// (UMMA M = 128); the three weight matrices stay resident in shared memory (bf16, K-major, 8x16-byte core matrices, no
// swizzle) for the whole launch; `tcgen05.mma.cta_group::1.kind::f16` accumulates each layer in fp32 in tensor memory
// (TMEM); every thread owns one row: it reads its accumulator row back with `tcgen05.ld`, adds the bias, applies tanh and
// writes the activations straight into shared memory as the next layer's A operand, so the hidden activations never touch
// HBM and there is one launch instead of nine.  HBM traffic: 72 B in, 16 B out per row.
//
// Numerics: layer 1 (whose inputs are raw observations of magnitude up to ~10) runs in TF32 (kind::tf32, 11-bit
// significands), layers 2 and 3 (inputs in [-1, 1]) in bf16; fp32 accumulation and bias everywhere, tanh.approx.f32.  Mean
// |out - fp32| ~ 2.5e-3; a trained policy has steep regions where any rounding moves an output by ~0.1, so the check that
// matters is behavioural: same win rate as the fp32 module (tests/test_actor_kernel.py).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace hk_actor {

constexpr int kRows = 128;   // rows per tile = UMMA M
constexpr int kObs = 18;     // observation width
constexpr int kK1 = 24;      // layer-1 K, padded to three TF32 UMMA K-steps of 8
constexpr int kHidden = 256;
constexpr int kN3 = 16;      // layer-3 N, padded from 4 to the smallest UMMA N for M = 128
constexpr int kAct = 4;
constexpr int kThreads = 128;

// packed parameter block (prepared once by the host side, see actor.py FusedActor): byte offsets
constexpr int kOffW1 = 0;                                // [kHidden x kK1]     tf32 (f32 bits), canonical K-major core-matrix order
constexpr int kOffW2 = kOffW1 + kHidden * kK1 * 4;       // [kHidden x kHidden] bf16
constexpr int kOffW3 = kOffW2 + kHidden * kHidden * 2;   // [kN3 x kHidden]     bf16 (rows 4..15 zero)
constexpr int kOffB1 = kOffW3 + kN3 * kHidden * 2;       // kHidden f32
constexpr int kOffB2 = kOffB1 + kHidden * 4;             // kHidden f32
constexpr int kOffB3 = kOffB2 + kHidden * 4;             // 16 f32 (4 used)
constexpr int kParamBytes = kOffB3 + 64;

// shared memory: the parameter block, then the A operands
constexpr int kOffA = kParamBytes;                       // [kRows x kHidden] bf16: layer-2 / layer-3 A operand
constexpr int kOffX = kOffA;                             // [kRows x kK1] tf32: layer-1 A operand (dead once layer 1 is done: aliased)
constexpr int kOffBar = kOffA + kRows * kHidden * 2;     // mbarrier (8 B) + TMEM base address (4 B)
constexpr int kSmemBytes = kOffBar + 16;

__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// UMMA shared-memory matrix descriptor, K-major, SWIZZLE_NONE: 8-row x 16-byte core matrices; `sbo` = byte distance between
// core matrices along M/N (8-row groups), `lbo` = byte distance between the two 16-byte K chunks of one K = 16 step
__device__ __forceinline__ uint64_t umaDesc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// instruction descriptor: D = f32, A = B = bf16 (fmt 1) or tf32 (fmt 2), both K-major, M x N
__host__ __device__ constexpr uint32_t instrDesc(int m, int n, uint32_t fmt) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmemD, uint64_t descA, uint64_t descB, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmemD),
      "l"(descA), "l"(descB), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void ummaTf32(uint32_t tmemD, uint64_t descA, uint64_t descB, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmemD),
      "l"(descA), "l"(descB), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ uint32_t toTf32(float x) {
  uint32_t y;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void ummaCommit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbarWait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
  } while (!ok);
}
__device__ __forceinline__ float tanhApprox(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t packBf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// 32 consecutive accumulator columns of this thread's TMEM lane (= its row)
__device__ __forceinline__ void tmemLoad32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
      "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(r[k]);
}

// hidden layer epilogue: this thread's 256 accumulator columns -> tanh(acc + bias) -> bf16 -> A operand row in smem
__device__ __forceinline__ void hiddenEpilogue(uint32_t taddr, const float* bias, unsigned char* sA) {
  const int t = threadIdx.x;
#pragma unroll 1
  for (int c = 0; c < kHidden / 32; ++c) {
    float v[32];
    tmemLoad32(taddr + (uint32_t)(c * 32), v);
#pragma unroll
    for (int q = 0; q < 4; ++q) {  // four 16-byte K chunks of 8 activations each
      uint32_t w[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int k = q * 8 + 2 * p;
        w[p] = packBf16(tanhApprox(v[k] + bias[c * 32 + k]), tanhApprox(v[k + 1] + bias[c * 32 + k + 1]));
      }
      // element (row t, column kc * 8 ..): chunk kc starts at kc * (kRows * 16) bytes, row t at + t * 16
      *reinterpret_cast<uint4*>(sA + (c * 4 + q) * (kRows * 16) + t * 16) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1) k_actor_mlp(const unsigned char* __restrict__ params, const float* __restrict__ obs,
                                                           float* __restrict__ act, int actStride, long long n) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int t = threadIdx.x, warp = t >> 5;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint32_t* tmemSlot = reinterpret_cast<uint32_t*>(smem + kOffBar + 8);
  // parameters: global (L2-resident after the first CTA) -> shared, once per CTA
  for (int k = t; k < kParamBytes / 16; k += kThreads)
    reinterpret_cast<uint4*>(smem)[k] = reinterpret_cast<const uint4*>(params)[k];
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smemAddr(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {  // one warp allocates all 512 TMEM columns: layer-1 accumulator [0,256), layer-2 [256,512), layer-3 [0,16)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smemAddr(tmemSlot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmemSlot;
  const uint32_t myT = tmem + ((uint32_t)(warp * 32) << 16);  // this warp's 32 TMEM lanes
  const uint32_t barA = smemAddr(bar);
  const uint32_t sW1 = smemAddr(smem + kOffW1), sW2 = smemAddr(smem + kOffW2), sW3 = smemAddr(smem + kOffW3);
  const uint32_t sX = smemAddr(smem + kOffX), sAa = smemAddr(smem + kOffA);
  const float* b1 = reinterpret_cast<const float*>(smem + kOffB1);
  const float* b2 = reinterpret_cast<const float*>(smem + kOffB2);
  const float* b3 = reinterpret_cast<const float*>(smem + kOffB3);
  unsigned char* sA = smem + kOffA;
  uint32_t phase = 0;
  const long long tiles = (n + kRows - 1) / kRows;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long row = tile * kRows + t;
    // ---- this thread's observation row -> tf32 -> layer-1 A operand (K padded 18 -> 24 with zeros) ----
    float o[kK1];
#pragma unroll
    for (int k = 0; k < kK1; ++k) o[k] = 0.0f;
    if (row < n) {
      const float2* src = reinterpret_cast<const float2*>(obs + row * kObs);
#pragma unroll
      for (int k = 0; k < kObs / 2; ++k) {
        const float2 v = src[k];
        o[2 * k] = v.x;
        o[2 * k + 1] = v.y;
      }
    }
#pragma unroll
    for (int q = 0; q < kK1 / 4; ++q)  // 16-byte K chunks of four tf32 values: chunk q at q * (kRows * 16) bytes, row t at + t * 16
      *reinterpret_cast<uint4*>(smem + kOffX + q * (kRows * 16) + t * 16) =
          make_uint4(toTf32(o[4 * q]), toTf32(o[4 * q + 1]), toTf32(o[4 * q + 2]), toTf32(o[4 * q + 3]));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    // ---- layer 1 (TF32): D1[128 x 256] = X[128 x 24] W1^T ----
    if (t == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int ks = 0; ks < kK1 / 8; ++ks)
        ummaTf32(tmem, umaDesc(sX + ks * 2 * (kRows * 16), kRows * 16, 128), umaDesc(sW1 + ks * 2 * (kHidden * 16), kHidden * 16, 128),
                 instrDesc(kRows, kHidden, 2), ks > 0);
      ummaCommit(barA);
    }
    mbarWait(barA, phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    hiddenEpilogue(myT, b1, sA);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    // ---- layer 2: D2[128 x 256] = H1[128 x 256] W2^T ----
    if (t == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int ks = 0; ks < kHidden / 16; ++ks)
        umma(tmem + kHidden, umaDesc(sAa + ks * 2 * (kRows * 16), kRows * 16, 128),
             umaDesc(sW2 + ks * 2 * (kHidden * 16), kHidden * 16, 128), instrDesc(kRows, kHidden, 1), ks > 0);
      ummaCommit(barA);
    }
    mbarWait(barA, phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    hiddenEpilogue(myT + kHidden, b2, sA);  // layer 2 has finished reading H1: its buffer takes H2
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    // ---- layer 3: D3[128 x 16] = H2[128 x 256] W3^T (4 real outputs) ----
    if (t == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int ks = 0; ks < kHidden / 16; ++ks)
        umma(tmem, umaDesc(sAa + ks * 2 * (kRows * 16), kRows * 16, 128), umaDesc(sW3 + ks * 2 * (kN3 * 16), kN3 * 16, 128),
             instrDesc(kRows, kN3, 1), ks > 0);
      ummaCommit(barA);
    }
    mbarWait(barA, phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
      uint32_t r0, r1, r2, r3;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(myT));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (row < n) {
        float* dst = act + row * actStride;
        dst[0] = tanhApprox(__uint_as_float(r0) + b3[0]);
        dst[1] = tanhApprox(__uint_as_float(r1) + b3[1]);
        dst[2] = tanhApprox(__uint_as_float(r2) + b3[2]);
        dst[3] = tanhApprox(__uint_as_float(r3) + b3[3]);
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();  // the next tile's layer 1 overwrites the columns layer 3 was just read from
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

}  // namespace hk_actor
