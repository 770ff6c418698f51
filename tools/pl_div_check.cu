// pl_div_check.cu -- is RN(a / s) reproduced by one correctly rounded reciprocal and Markstein's correction step?
// (field_pl_kernel shares the reciprocal of a sample's normalisation sum between its three quotients,
// /root/reference/bcf_filtered_reader.cpp:276-287.)  Compares bit for bit against __ddiv_rn:
//   1. all 256^3 phred triples under the first EM iteration's weights (af = 1/2) and under skewed weights
//   2. 2^30 random (weights, phred triple) operands with allele frequencies down to 1e-9
//   3. 2^30 quotients a / s with a <= s anywhere in [2^-61, 1] (the entry-table kernels' normalisations)
// prints one JSON line; exit code 1 on any mismatch.  nvcc -arch=sm_100a -O3 -o pl_div_check pl_div_check.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ double div_shared(double a, double s, double r) {
  const double q = __dmul_rn(a, r);
  const double rem = __fma_rn(-q, s, a);
  return __fma_rn(rem, r, q);
}

__device__ unsigned long long g_bad, g_n;

__device__ __forceinline__ void check3(double g0, double g1, double g2) {
  const double sum = __dadd_rn(__dadd_rn(__dadd_rn(0.0, g0), g1), g2);
  const double rs = __drcp_rn(sum);
  const double a[3] = {g0, g1, g2};
  unsigned bad = 0;
#pragma unroll
  for (int i = 0; i < 3; ++i)
    bad += __double_as_longlong(div_shared(a[i], sum, rs)) != __double_as_longlong(__ddiv_rn(a[i], sum));
  if (bad) atomicAdd(&g_bad, (unsigned long long)bad);
}

__global__ void exhaustive(const double* __restrict__ prob, double q0, double q1, double q2) {
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;   // 2^24 triples
  const unsigned a = t & 255, b = (t >> 8) & 255, c = t >> 16;
  check3(__dmul_rn(q0, prob[a]), __dmul_rn(q1, prob[b]), __dmul_rn(q2, prob[c]));
  if (t == 0) atomicAdd(&g_n, 3ull << 24);
}

__global__ void randomised(const double* __restrict__ prob, unsigned long long seed, int per_thread) {
  unsigned long long s = seed + (unsigned long long)(blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
  for (int i = 0; i < per_thread; ++i) {
    const unsigned long long r = rnd();
    // allele frequency log-uniform in [1e-9, 1 - 1e-9] or any double in (0, 1)
    double af = (r & 1) ? exp(-20.7 * (double)(rnd() >> 11) * (1.0 / 9007199254740992.0))
                        : (double)(rnd() >> 11) * (1.0 / 9007199254740992.0);
    if (!(af > 0.0) || af >= 1.0) af = 0.5;
    const double bf = 1.0 - af;
    const unsigned a = (r >> 8) & 255, b = (r >> 16) & 255, c = (r >> 24) & 255;
    check3(__dmul_rn(__dmul_rn(af, af), prob[a]), __dmul_rn(__dmul_rn(__dmul_rn(2.0, bf), af), prob[b]),
           __dmul_rn(__dmul_rn(bf, bf), prob[c]));
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&g_n, 3ull * gridDim.x * blockDim.x * per_thread);
}

// 3. the entry-table kernels' use (entq1_kernel: P[v] / max over v): a <= s anywhere in [2^-60, 1]
__global__ void generic(unsigned long long seed, int per_thread) {
  unsigned long long s = seed + (unsigned long long)(blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
  unsigned long long bad = 0;
  for (int i = 0; i < per_thread; ++i) {
    const unsigned long long r = rnd();
    const double sv = ldexp(1.0 + (double)(rnd() >> 12) * (1.0 / 4503599627370496.0), -(int)(r % 61) - 1);
    const double av = sv * ((double)(rnd() >> 11) * (1.0 / 9007199254740992.0));
    const double rs = __drcp_rn(sv);
    bad += __double_as_longlong(div_shared(av, sv, rs)) != __double_as_longlong(__ddiv_rn(av, sv));
    bad += __double_as_longlong(div_shared(sv, sv, rs)) != __double_as_longlong(__ddiv_rn(sv, sv));
  }
  if (bad) atomicAdd(&g_bad, bad);
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&g_n, 2ull * gridDim.x * blockDim.x * per_thread);
}

int main() {
  double h[256];
  for (int i = 0; i < 256; ++i) h[i] = pow(0.1, i * 0.1);   // PhredHelper.cpp:31
  double* d;
  cudaMalloc(&d, sizeof(h));
  cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
  unsigned long long z = 0;
  cudaMemcpyToSymbol(g_bad, &z, 8);
  cudaMemcpyToSymbol(g_n, &z, 8);
  const double afs[][2] = {{0.5, 0.5}, {0.9, 0.1}, {0.999, 0.001}, {1.0 - 1e-7, 1e-7}, {0.3, 0.7}, {1e-5, 1.0 - 1e-5}};
  for (auto& f : afs) exhaustive<<<1 << 16, 256>>>(d, f[0] * f[0], 2.0 * f[1] * f[0], f[1] * f[1]);
  randomised<<<148 * 16, 256>>>(d, 0x1234567ull, 1772);     // ~2^30 triples
  generic<<<148 * 16, 256>>>(0x7654321ull, 886);            // ~2^30 quotients
  cudaDeviceSynchronize();
  unsigned long long bad = 0, n = 0;
  cudaMemcpyFromSymbol(&bad, g_bad, 8);
  cudaMemcpyFromSymbol(&n, g_n, 8);
  printf("{\"check\": \"RN(a/s) from one reciprocal + Markstein step vs __ddiv_rn\", \"quotients\": %llu, \"mismatches\": %llu, "
         "\"cuda_error\": \"%s\"}\n", n, bad, cudaGetErrorString(cudaGetLastError()));
  return bad != 0;
}
