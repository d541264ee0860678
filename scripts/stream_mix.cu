// Micro-benchmark (diagnostic, not product code): what HBM rate can a kernel with the sweep kernels' stream mix reach?
// NIN input vectors are streamed in, NOUT output vectors streamed out (out_k = in_0 + c_k * in_k, no stencil), with
//   mode 0: plain vectorised loads (ld.global.nc double2) and st.global.cs stores, grid-stride
//   mode 1: the product's structure - producer warp + cp.async.bulk rows into an mbarrier stage ring, consumer warps
//           store with st.global.cs
//   mode 2: as 1, but consumers write their results back into the stage (in place) and one thread emits them with
//           cp.async.bulk shared -> global row stores (bulk_group completion gates the slot's reuse)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/stream_mix scripts/stream_mix.cu
// Run:   scripts/stream_mix [elements per vector, default 2^28]
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e = (x);                                                               \
    if (e != cudaSuccess) {                                                            \
      printf("%s failed: %s (line %d)\n", #x, cudaGetErrorString(e), __LINE__);        \
      exit(1);                                                                         \
    }                                                                                  \
  } while (0)

constexpr int ROW = 512;  // doubles per row (4 KB)
constexpr int CONS_WARPS = 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(
          smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

struct Args {
  const double* in[3];
  double* out[3];
  size_t rows;  // rows of ROW doubles per vector
};

template <int NIN, int NOUT>
__global__ void __launch_bounds__(256) plain_kernel(Args a) {
  const size_t n2 = a.rows * ROW / 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
    double2 v[NIN];
#pragma unroll
    for (int k = 0; k < NIN; ++k) v[k] = __ldg(reinterpret_cast<const double2*>(a.in[k]) + i);
#pragma unroll
    for (int k = 0; k < NOUT; ++k) {
      double2 o;
      o.x = v[0].x + 1.0000001 * v[k < NIN ? k : 0].x;
      o.y = v[0].y + 1.0000001 * v[k < NIN ? k : 0].y;
      __stcs(reinterpret_cast<double2*>(a.out[k]) + i, o);
    }
  }
}

// MODE 1: bulk loads, STG stores.  MODE 2: bulk loads, in-place results, bulk stores.
template <int NIN, int NOUT, int HS, int NST, int MODE>
__global__ void __launch_bounds__((CONS_WARPS + 2) * 32, 2) ring_kernel(Args a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int STAGE = HS * NIN * ROW;
  double* data = reinterpret_cast<double*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NST * STAGE * 8);
  uint64_t* empty = full + NST;
  uint64_t* ready = empty + NST;  // MODE 2: results of a stage are in shared memory
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int i = 0; i < NST; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], MODE == 2 ? 1 : CONS_WARPS);
      mbar_init(&ready[i], CONS_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // contiguous chunk of rows per CTA
  const size_t per = (a.rows + gridDim.x - 1) / gridDim.x;
  const size_t r0 = (size_t)blockIdx.x * per, r1 = r0 + per < a.rows ? r0 + per : a.rows;
  if (warp == CONS_WARPS) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (size_t r = r0; r < r1; r += HS) {
        mbar_wait(&empty[stage], phase ^ 1u);
        const int nr = (int)(r1 - r < (size_t)HS ? r1 - r : HS);
        mbar_arrive_expect_tx(&full[stage], (uint32_t)(nr * NIN * ROW * 8));
        double* sd = data + (size_t)stage * STAGE;
        for (int j = 0; j < nr; ++j)
          for (int k = 0; k < NIN; ++k) bulk_g2s(sd + (k * HS + j) * ROW, a.in[k] + (r + j) * ROW, ROW * 8, &full[stage]);
        if (++stage == NST) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == CONS_WARPS + 1) {
    // store warp (MODE 2): emits a finished stage with bulk shared -> global row copies, then frees the slot
    if (MODE == 2 && lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (size_t r = r0; r < r1; r += HS) {
        mbar_wait(&ready[stage], phase);
        const int nr = (int)(r1 - r < (size_t)HS ? r1 - r : HS);
        double* sd = data + (size_t)stage * STAGE;
        for (int j = 0; j < nr; ++j)
          for (int k = 0; k < NOUT; ++k) bulk_s2g(a.out[k] + (r + j) * ROW, sd + (k * HS + j) * ROW, ROW * 8);
        bulk_commit();
        bulk_wait_read<0>();  // the stores have read the slot: hand it back
        mbar_arrive(&empty[stage]);
        if (++stage == NST) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    int stage = 0;
    uint32_t phase = 0;
    for (size_t r = r0; r < r1; r += HS) {
      mbar_wait(&full[stage], phase);
      const int nr = (int)(r1 - r < (size_t)HS ? r1 - r : HS);
      double* sd = data + (size_t)stage * STAGE;
      for (int j = 0; j < nr; ++j) {
        double2 v[NIN];
#pragma unroll
        for (int k = 0; k < NIN; ++k) v[k] = *reinterpret_cast<const double2*>(sd + (k * HS + j) * ROW + 2 * tid);
#pragma unroll
        for (int k = 0; k < NOUT; ++k) {
          double2 o;
          o.x = v[0].x + 1.0000001 * v[k < NIN ? k : 0].x;
          o.y = v[0].y + 1.0000001 * v[k < NIN ? k : 0].y;
          if (MODE == 1) __stcs(reinterpret_cast<double2*>(a.out[k] + (r + j) * ROW + 2 * tid), o);
          else *reinterpret_cast<double2*>(sd + (k * HS + j) * ROW + 2 * tid) = o;  // NOUT <= NIN: in place
        }
      }
      if (MODE == 1) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
      } else {
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ready[stage]);
      }
      if (++stage == NST) { stage = 0; phase ^= 1u; }
    }
  }
}

// The product's access pattern: the vectors are 2-D [R][pitch]; a CTA marches down a strip of COLS columns, so its
// consecutive row copies are `pitch` doubles apart in memory (131 KB at 16384^2) instead of adjacent.
template <int NIN, int NOUT, int HS, int NST, int COLS, int SPLIT = 0>
__global__ void __launch_bounds__((CONS_WARPS + 1) * 32, 2) strip_kernel(Args a, int pitch, int R) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int STAGE = HS * NIN * COLS;
  double* data = reinterpret_cast<double*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NST * STAGE * 8);
  uint64_t* empty = full + NST;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int i = 0; i < NST; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], CONS_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int strips = pitch / COLS;
  const long long total = (long long)strips * R;  // (strip, row) pairs, strip-major
  const long long per = (total + gridDim.x - 1) / gridDim.x;
  const long long t0 = (long long)blockIdx.x * per, t1 = t0 + per < total ? t0 + per : total;
  if (warp == CONS_WARPS) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long t = t0; t < t1;) {
        const int strip = (int)(t / R), row = (int)(t % R);
        const int nr = (int)min((long long)HS, min((long long)(R - row), t1 - t));
        mbar_wait(&empty[stage], phase ^ 1u);
        mbar_arrive_expect_tx(&full[stage], (uint32_t)(nr * NIN * COLS * 8));
        double* sd = data + (size_t)stage * STAGE;
        for (int j = 0; j < nr; ++j)
          for (int k = 0; k < NIN; ++k) {
            double* dst = sd + (k * HS + j) * COLS;
            const double* src = a.in[k] + (size_t)(row + j) * pitch + (size_t)strip * COLS;
            if (SPLIT) {  // head up to the next 128-byte line, aligned middle, tail
              const int head = (int)((16 - (((size_t)strip * COLS) & 15)) & 15);
              const int mid = (COLS - head) & ~15, tail = COLS - head - mid;
              if (head) bulk_g2s(dst, src, head * 8, &full[stage]);
              bulk_g2s(dst + head, src + head, mid * 8, &full[stage]);
              if (tail) bulk_g2s(dst + head + mid, src + head + mid, tail * 8, &full[stage]);
            } else {
              bulk_g2s(dst, src, COLS * 8, &full[stage]);
            }
          }
        t += nr;
        if (++stage == NST) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    int stage = 0;
    uint32_t phase = 0;
    for (long long t = t0; t < t1;) {
      const int strip = (int)(t / R), row = (int)(t % R);
      const int nr = (int)min((long long)HS, min((long long)(R - row), t1 - t));
      mbar_wait(&full[stage], phase);
      const double* sd = data + (size_t)stage * STAGE;
      for (int c = 2 * tid; c + 1 < COLS; c += 2 * CONS_WARPS * 32) {
        for (int j = 0; j < nr; ++j) {
          double2 v[NIN];
#pragma unroll
          for (int k = 0; k < NIN; ++k) v[k] = *reinterpret_cast<const double2*>(sd + (k * HS + j) * COLS + c);
#pragma unroll
          for (int k = 0; k < NOUT; ++k) {
            double2 o;
            o.x = v[0].x + 1.0000001 * v[k < NIN ? k : 0].x;
            o.y = v[0].y + 1.0000001 * v[k < NIN ? k : 0].y;
            __stcs(reinterpret_cast<double2*>(a.out[k] + (size_t)(row + j) * pitch + (size_t)strip * COLS + c), o);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
      t += nr;
      if (++stage == NST) { stage = 0; phase ^= 1u; }
    }
  }
}

// The product's geometry: a strip OWNS `OWN` columns (stores exactly those) and stages LOAD >= OWN columns starting
// SHIFT columns to the left (halo columns); throughput is counted on the owned bytes only.
template <int NIN, int NOUT, int HS, int NST, int OWN, int LOAD, int SHIFT>
__global__ void __launch_bounds__((CONS_WARPS + 1) * 32, 2) halo_kernel(Args a, int pitch, int R) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int STAGE = HS * NIN * LOAD;
  double* data = reinterpret_cast<double*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NST * STAGE * 8);
  uint64_t* empty = full + NST;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int i = 0; i < NST; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], CONS_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int strips = (pitch - 64) / OWN;  // (the first strip starts 32 columns in, so halos stay inside the row)
  const long long total = (long long)strips * R;
  const long long per = (total + gridDim.x - 1) / gridDim.x;
  const long long t0 = (long long)blockIdx.x * per, t1 = t0 + per < total ? t0 + per : total;
  if (warp == CONS_WARPS) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long t = t0; t < t1;) {
        const int strip = (int)(t / R), row = (int)(t % R);
        const int nr = (int)min((long long)HS, min((long long)(R - row), t1 - t));
        mbar_wait(&empty[stage], phase ^ 1u);
        mbar_arrive_expect_tx(&full[stage], (uint32_t)(nr * NIN * LOAD * 8));
        double* sd = data + (size_t)stage * STAGE;
        for (int j = 0; j < nr; ++j)
          for (int k = 0; k < NIN; ++k)
            bulk_g2s(sd + (k * HS + j) * LOAD, a.in[k] + (size_t)(row + j) * pitch + 32 + (size_t)strip * OWN - SHIFT, LOAD * 8,
                     &full[stage]);
        t += nr;
        if (++stage == NST) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    int stage = 0;
    uint32_t phase = 0;
    for (long long t = t0; t < t1;) {
      const int strip = (int)(t / R), row = (int)(t % R);
      const int nr = (int)min((long long)HS, min((long long)(R - row), t1 - t));
      mbar_wait(&full[stage], phase);
      const double* sd = data + (size_t)stage * STAGE;
      for (int c = 2 * tid; c + 1 < OWN; c += 2 * CONS_WARPS * 32) {
        for (int j = 0; j < nr; ++j) {
          double2 v[NIN];
#pragma unroll
          for (int k = 0; k < NIN; ++k) v[k] = *reinterpret_cast<const double2*>(sd + (k * HS + j) * LOAD + SHIFT + c);
#pragma unroll
          for (int k = 0; k < NOUT; ++k) {
            double2 o;
            o.x = v[0].x + 1.0000001 * v[k < NIN ? k : 0].x;
            o.y = v[0].y + 1.0000001 * v[k < NIN ? k : 0].y;
            __stcs(reinterpret_cast<double2*>(a.out[k] + (size_t)(row + j) * pitch + 32 + (size_t)strip * OWN + c), o);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
      t += nr;
      if (++stage == NST) { stage = 0; phase ^= 1u; }
    }
  }
}
template <int NIN, int NOUT, int HS, int NST, int OWN, int LOAD, int SHIFT>
float run_halo(const Args& a, int sms, int reps, int pitch, int R) {
  auto k = halo_kernel<NIN, NOUT, HS, NST, OWN, LOAD, SHIFT>;
  const size_t smem = (size_t)NST * HS * NIN * LOAD * 8 + NST * 16 + 128;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) k<<<2 * sms, (CONS_WARPS + 1) * 32, smem>>>(a, pitch, R);
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) k<<<2 * sms, (CONS_WARPS + 1) * 32, smem>>>(a, pitch, R);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / reps;
}

template <int NIN, int NOUT, int HS, int NST, int COLS, int SPLIT = 0>
float run_strip(const Args& a, int sms, int reps, int pitch, int R) {
  auto k = strip_kernel<NIN, NOUT, HS, NST, COLS, SPLIT>;
  const size_t smem = (size_t)NST * HS * NIN * COLS * 8 + NST * 16 + 128;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) k<<<2 * sms, (CONS_WARPS + 1) * 32, smem>>>(a, pitch, R);
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) k<<<2 * sms, (CONS_WARPS + 1) * 32, smem>>>(a, pitch, R);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / reps;
}

template <int NIN, int NOUT, int HS, int NST, int MODE>
float run_ring(const Args& a, int sms, int reps) {
  auto k = ring_kernel<NIN, NOUT, HS, NST, MODE>;
  const size_t smem = (size_t)NST * HS * NIN * ROW * 8 + NST * 24 + 128;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) k<<<2 * sms, (CONS_WARPS + 2) * 32, smem>>>(a);
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) k<<<2 * sms, (CONS_WARPS + 2) * 32, smem>>>(a);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / reps;
}
template <int NIN, int NOUT>
float run_plain(const Args& a, int sms, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) plain_kernel<NIN, NOUT><<<sms * 8, 256>>>(a);
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) plain_kernel<NIN, NOUT><<<sms * 8, 256>>>(a);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / reps;
}

int main(int argc, char** argv) {
  const size_t n = argc > 1 ? strtoull(argv[1], nullptr, 10) : (1ull << 28);  // 2 GiB per vector
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  Args a;
  a.rows = n / ROW;
  for (int k = 0; k < 3; ++k) {
    CK(cudaMalloc((void**)&a.in[k], n * 8));
    CK(cudaMalloc((void**)&a.out[k], n * 8));
    CK(cudaMemset((void*)a.in[k], 0, n * 8));
    CK(cudaMemset(a.out[k], 0, n * 8));
  }
  const int reps = 20;
  auto report = [&](const char* name, int nin, int nout, float ms) {
    printf("%-44s %d in %d out  %8.3f ms  %8.1f GB/s\n", name, nin, nout, ms, (double)(nin + nout) * n * 8 / ms / 1e6);
    fflush(stdout);
  };
  {  // cudaMemcpy D2D as the reference point
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) CK(cudaMemcpyAsync(a.out[0], a.in[0], n * 8, cudaMemcpyDeviceToDevice));
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) CK(cudaMemcpyAsync(a.out[0], a.in[0], n * 8, cudaMemcpyDeviceToDevice));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    report("cudaMemcpy D2D", 1, 1, ms / reps);
  }
  report("plain ld/st.cs", 1, 0, run_plain<1, 0>(a, sms, reps));
  report("plain ld/st.cs", 2, 0, run_plain<2, 0>(a, sms, reps));
  report("plain ld/st.cs", 1, 1, run_plain<1, 1>(a, sms, reps));
  report("plain ld/st.cs", 2, 2, run_plain<2, 2>(a, sms, reps));
  report("plain ld/st.cs", 3, 3, run_plain<3, 3>(a, sms, reps));
  report("bulk-load ring HS4 + st.cs", 2, 0, run_ring<2, 0, 4, 3, 1>(a, sms, reps));
  report("bulk-load ring HS4 + st.cs", 1, 1, run_ring<1, 1, 4, 6, 1>(a, sms, reps));
  report("bulk-load ring HS4 + st.cs", 2, 2, run_ring<2, 2, 4, 3, 1>(a, sms, reps));
  report("bulk-load ring HS4 + st.cs", 3, 3, run_ring<3, 3, 4, 2, 1>(a, sms, reps));
  report("bulk-load ring HS2 + st.cs", 2, 2, run_ring<2, 2, 2, 6, 1>(a, sms, reps));
  report("bulk-load ring HS2 + st.cs", 3, 3, run_ring<3, 3, 2, 4, 1>(a, sms, reps));
  report("bulk-load ring HS4 + bulk store (in place)", 1, 1, run_ring<1, 1, 4, 6, 2>(a, sms, reps));
  report("bulk-load ring HS4 + bulk store (in place)", 2, 2, run_ring<2, 2, 4, 3, 2>(a, sms, reps));
  report("bulk-load ring HS4 + bulk store (in place)", 3, 3, run_ring<3, 3, 4, 2, 2>(a, sms, reps));
  report("bulk-load ring HS2 + bulk store (in place)", 2, 2, run_ring<2, 2, 2, 6, 2>(a, sms, reps));
  report("bulk-load ring HS2 + bulk store (in place)", 3, 3, run_ring<3, 3, 2, 4, 2>(a, sms, reps));
  {  // strips of a pitched 2-D array (16384 rows x 16384 columns: 2 GiB per vector)
    const int pitch = 16384, R = (int)(n / pitch);
    auto rep2 = [&](const char* name, int nin, int nout, int cols, float ms) {
      const double bytes = (double)(nin + nout) * (double)(pitch / cols) * cols * R * 8;
      printf("%-44s %d in %d out  %8.3f ms  %8.1f GB/s\n", name, nin, nout, ms, bytes / ms / 1e6);
      fflush(stdout);
    };
    rep2("strip march, 512-column strips, HS4", 2, 2, 512, run_strip<2, 2, 4, 3, 512>(a, sms, reps, pitch, R));
    rep2("strip march, 512-column strips, HS4", 3, 3, 512, run_strip<3, 3, 4, 2, 512>(a, sms, reps, pitch, R));
    rep2("strip march, 424-column strips, HS4", 2, 2, 424, run_strip<2, 2, 4, 3, 424>(a, sms, reps, pitch, R));
    rep2("strip march, 424-column strips, HS4", 3, 3, 424, run_strip<3, 3, 4, 2, 424>(a, sms, reps, pitch, R));
    rep2("strip march, 424 columns, split at 128-B lines", 2, 2, 424, run_strip<2, 2, 4, 3, 424, 1>(a, sms, reps, pitch, R));
    rep2("strip march, 424 columns, split at 128-B lines", 3, 3, 424, run_strip<3, 3, 4, 2, 424, 1>(a, sms, reps, pitch, R));
    rep2("strip march, 416-column strips, HS4", 2, 2, 416, run_strip<2, 2, 4, 3, 416>(a, sms, reps, pitch, R));
    rep2("strip march, 432-column strips, HS4", 2, 2, 432, run_strip<2, 2, 4, 3, 432>(a, sms, reps, pitch, R));
    rep2("strip march, 448-column strips, HS4", 2, 2, 448, run_strip<2, 2, 4, 3, 448>(a, sms, reps, pitch, R));
    rep2("strip march, 448-column strips, HS4", 3, 3, 448, run_strip<3, 3, 4, 2, 448>(a, sms, reps, pitch, R));
    auto rep3 = [&](const char* name, int nin, int nout, int own, float ms) {
      const double bytes = (double)(nin + nout) * (double)((pitch - 64) / own) * own * R * 8;
      printf("%-52s %d in %d out  %8.3f ms  %8.1f GB/s of OWNED bytes\n", name, nin, nout, ms, bytes / ms / 1e6);
      fflush(stdout);
    };
    rep3("own 420 (32-B aligned), load 424 @-2 [product]", 2, 2, 420, run_halo<2, 2, 4, 3, 420, 424, 2>(a, sms, reps, pitch, R));
    rep3("own 420 (32-B aligned), load 424 @-2 [product]", 3, 3, 420, run_halo<3, 3, 4, 2, 420, 424, 2>(a, sms, reps, pitch, R));
    rep3("own 416 (line aligned), load 420 @-2", 2, 2, 416, run_halo<2, 2, 4, 3, 416, 420, 2>(a, sms, reps, pitch, R));
    rep3("own 416 (line aligned), load 420 @-2", 3, 3, 416, run_halo<3, 3, 4, 2, 416, 420, 2>(a, sms, reps, pitch, R));
    rep3("own 416 (line aligned), load 448 @-16 (aligned)", 2, 2, 416, run_halo<2, 2, 4, 3, 416, 448, 16>(a, sms, reps, pitch, R));
    rep3("own 416 (line aligned), load 448 @-16 (aligned)", 3, 3, 416, run_halo<3, 3, 4, 2, 416, 448, 16>(a, sms, reps, pitch, R));
    rep3("own 832 (line aligned), load 864 @-16 (aligned)", 2, 2, 832, run_halo<2, 2, 2, 3, 832, 864, 16>(a, sms, reps, pitch, R));
    rep3("own 480 (line aligned), load 484 @-2", 2, 2, 480, run_halo<2, 2, 4, 3, 480, 484, 2>(a, sms, reps, pitch, R));
    rep2("strip march, 256-column strips, HS4", 2, 2, 256, run_strip<2, 2, 4, 6, 256>(a, sms, reps, pitch, R));
    rep2("strip march, 1024-column strips, HS2", 2, 2, 1024, run_strip<2, 2, 2, 3, 1024>(a, sms, reps, pitch, R));
  }
  return 0;
}
