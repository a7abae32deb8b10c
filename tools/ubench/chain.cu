// Micro-benchmark (developer tool): latency of the DP cost-chain variants on one warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false chain.cu -o chain && ./chain
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ double round_int(double x) {
  long long b = __double_as_longlong(x);
  b += 0x0FFFFFFFLL + ((b >> 29) & 1);
  b &= ~0x1FFFFFFFLL;
  return __longlong_as_double(b);
}
__device__ __forceinline__ double round_cvt(double x) { return (double)(float)x; }

template <int MODE>
__global__ void k(double* out, const double* in, int n, long long* cyc) {
  double cj = in[0];
  const double llb = in[1 + (threadIdx.x & 3)], e2 = in[8];
  double acc = 0;
  long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < n; i++) {
    const double lit = llb + cj;
    double e = e2 + acc;  // something that varies
    if (MODE == 0) { const bool take = lit < e; cj = take ? round_int(lit) : e; }
    if (MODE == 1) { const bool take = lit < e; cj = take ? round_cvt(lit) : e; }
    if (MODE == 2) { cj = round_int(lit); }
    if (MODE == 3) { cj = round_cvt(lit); }
    if (MODE == 4) { cj = lit; }
    if (MODE == 5) { const bool take = lit < e; cj = take ? lit : e; }
    if (MODE == 6) {  // integer compare on the bit patterns (non-negative doubles)
      const bool take = __double_as_longlong(lit) < __double_as_longlong(e);
      cj = take ? round_int(lit) : e;
    }
    if (MODE == 7) {  // select on rounded values: min in the integer domain after rounding both
      const long long a = __double_as_longlong(round_int(lit)), b = __double_as_longlong(e);
      cj = __longlong_as_double(a < b ? a : b);
    }
    acc += 1e-9;
  }
  long long t1 = clock64();
  out[threadIdx.x] = cj + acc;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  double h[16] = {0.0, 3.25, 4.5, 5.125, 6.0625, 0, 0, 0, 1e30};
  double *din, *dout; long long* dc;
  cudaMalloc(&din, sizeof(h)); cudaMalloc(&dout, 32 * 8); cudaMalloc(&dc, 8);
  cudaMemcpy(din, h, sizeof(h), cudaMemcpyHostToDevice);
  const int n = 1 << 16;
  const char* names[] = {"dsetp+int round+sel", "dsetp+cvt round+sel", "int round only", "cvt round only", "dadd only",
                         "dsetp+sel no round", "isetp64+int round+sel", "int round + imin64"};
  for (int m = 0; m < 8; m++) {
    for (int rep = 0; rep < 2; rep++) {
      switch (m) {
        case 0: k<0><<<1, 32>>>(dout, din, n, dc); break;
        case 1: k<1><<<1, 32>>>(dout, din, n, dc); break;
        case 2: k<2><<<1, 32>>>(dout, din, n, dc); break;
        case 3: k<3><<<1, 32>>>(dout, din, n, dc); break;
        case 4: k<4><<<1, 32>>>(dout, din, n, dc); break;
        case 5: k<5><<<1, 32>>>(dout, din, n, dc); break;
        case 6: k<6><<<1, 32>>>(dout, din, n, dc); break;
        case 7: k<7><<<1, 32>>>(dout, din, n, dc); break;
      }
      cudaDeviceSynchronize();
    }
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s %.1f cycles/iter\n", names[m], (double)c / n);
  }
  return 0;
}
