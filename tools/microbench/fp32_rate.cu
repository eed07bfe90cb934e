// fp32_rate.cu -- issue rate of scalar FFMA / FMUL / FADD against the packed FFMA2 / FMUL2 / FADD2 forms
// (fma.rn.f32x2, sm_100+) and of FFMA mixed with LDS, per SM sub-partition.  Development probe, not part of
// the product: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_rate fp32_rate.cu && ./fp32_rate
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 2048, CH = 16;

template <int MODE> __global__ void __launch_bounds__(256) k(float *out, const float *in, long long *cyc) {
  __shared__ float sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = in[i];
  __syncthreads();
  float a[CH], b = in[threadIdx.x], c = in[threadIdx.x + 256];
  float2 a2[CH / 2], b2 = make_float2(b, c), c2 = make_float2(c, b);
#pragma unroll
  for (int i = 0; i < CH; i++) a[i] = in[i + threadIdx.x];
#pragma unroll
  for (int i = 0; i < CH / 2; i++) a2[i] = make_float2(a[2 * i], a[2 * i + 1]);
  int idx = threadIdx.x & 31;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; it++) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < CH; i++) a[i] = __fmaf_rn(a[i], b, c);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < CH / 2; i++) a2[i] = __ffma2_rn(a2[i], b2, c2);
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < CH; i++) a[i] = __fmul_rn(a[i], b);
    } else if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < CH; i++) a[i] = __fadd_rn(a[i], b);
    } else if (MODE == 4) {  // 16 FFMA + 8 LDS per iteration
#pragma unroll
      for (int i = 0; i < CH; i++) a[i] = __fmaf_rn(a[i], b, c);
#pragma unroll
      for (int i = 0; i < 8; i++) c += sm[(idx + 32 * i) & 1023];
      idx = (idx + 1) & 1023;
    } else if (MODE == 5) {  // 8 FFMA2 (= 16 FMAs) + 8 LDS
#pragma unroll
      for (int i = 0; i < CH / 2; i++) a2[i] = __ffma2_rn(a2[i], b2, c2);
#pragma unroll
      for (int i = 0; i < 8; i++) c2.x += sm[(idx + 32 * i) & 1023];
      idx = (idx + 1) & 1023;
    } else if (MODE == 6) {  // FFMA whose three sources are distinct registers (no operand reuse)
#pragma unroll
      for (int i = 0; i < CH; i++) a[i] = __fmaf_rn(a[(i + 1) % CH], a[(i + 5) % CH], a[i]);
    } else if (MODE == 8) {  // FFMA2 with one scalar operand broadcast to both halves (R.F32 operand form)
#pragma unroll
      for (int i = 0; i < CH / 2; i++) a2[i] = __ffma2_rn(a2[i], make_float2(b, b), c2);
    } else if (MODE == 9) {  // FMUL2 / FADD2
#pragma unroll
      for (int i = 0; i < CH / 2; i++) a2[i] = __fadd2_rn(__fmul2_rn(a2[i], b2), c2);
    } else if (MODE == 7) {  // 16 FFMA + 16 IADD (alu pipe)
#pragma unroll
      for (int i = 0; i < CH; i++) a[i] = __fmaf_rn(a[i], b, c);
#pragma unroll
      for (int i = 0; i < CH; i++) idx = (idx + (i + 1)) ^ it;
    }
  }
  const long long t1 = clock64();
  float s = c + c2.x + c2.y + (float)idx;
#pragma unroll
  for (int i = 0; i < CH; i++) s += a[i];
#pragma unroll
  for (int i = 0; i < CH / 2; i++) s += a2[i].x + a2[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char *name, double fma_per_thread_iter, int ctas_per_sm) {
  int nsm;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  const int nb = nsm * ctas_per_sm;
  float *out, *in;
  long long *cyc;
  cudaMalloc(&out, sizeof(float) * nb * 256);
  cudaMalloc(&in, sizeof(float) * 2048);
  cudaMemset(in, 0, sizeof(float) * 2048);
  cudaMalloc(&cyc, sizeof(long long) * nb);
  k<MODE><<<nb, 256>>>(out, in, cyc);
  k<MODE><<<nb, 256>>>(out, in, cyc);
  cudaDeviceSynchronize();
  long long h[4096];
  cudaMemcpy(h, cyc, sizeof(long long) * nb, cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < nb; i++) mean += (double)h[i];
  mean /= nb;
  // per SM: ctas_per_sm * 256 threads * ITER * ops / cycles
  const double per_clk_sm = ctas_per_sm * 256.0 * ITER * fma_per_thread_iter / mean;
  printf("%-34s ctas/SM %d  cycles %9.0f  fp32 ops/clk/SM %7.1f  (warp-instr/clk/SMSP %.3f)\n", name, ctas_per_sm, mean,
         per_clk_sm, per_clk_sm / 128.0);
  cudaFree(out); cudaFree(in); cudaFree(cyc);
}

int main() {
  for (int c : {1, 2, 4}) {
    run<0>("FFMA  (a = a*b + c)", CH, c);
    run<6>("FFMA  (three distinct sources)", CH, c);
    run<1>("FFMA2 (fma.rn.f32x2)", CH, c);
    run<8>("FFMA2 (scalar broadcast operand)", CH, c);
    run<9>("FMUL2 + FADD2", 2 * CH, c);
    run<2>("FMUL", CH, c);
    run<3>("FADD", CH, c);
    run<4>("16 FFMA + 8 LDS", CH, c);
    run<5>("8 FFMA2 + 8 LDS", CH, c);
    run<7>("16 FFMA + 16 integer ops", CH, c);
  }
  return 0;
}
