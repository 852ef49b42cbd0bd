// Latency and issue rate of the packed fp32x2 operations the ICP kernel is built on (FFMA2 / FMUL2 /
// FADD2) against scalar FFMA, per SM sub-partition, on the GPU it runs on.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/fp32x2_probe tools/fp32x2_probe.cu && /tmp/fp32x2_probe
// Output: cycles per warp-instruction for a dependent chain (latency) and for ILP independent chains
// with W warps on one sub-partition (throughput).
#include <cstdio>
#include <cuda_runtime.h>

constexpr int N = 4096;

template <int OP, int ILP>
__global__ void probe(float2* out, long long* cycles, float2 seed) {
  float2 x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = make_float2(seed.x + i, seed.y - i);
  const float2 m = make_float2(1.0000001f, 0.9999999f), c = make_float2(1e-7f, -1e-7f);
  float2 y[ILP], z[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) {
    y[i] = make_float2(1.0f + 1e-7f * (threadIdx.x + i), 1.0f - 1e-7f * (threadIdx.x + 2 * i));
    z[i] = make_float2(1e-7f * (threadIdx.x + 3 * i), -1e-7f * (threadIdx.x + i));
  }
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int k = 0; k < N; k += 8) {
#pragma unroll
    for (int u = 0; u < 8; u++)
#pragma unroll
      for (int i = 0; i < ILP; i++) {
        if (OP == 0) x[i].x = __fmaf_rn(x[i].x, m.x, c.x);
        if (OP == 1) x[i] = __ffma2_rn(x[i], m, c);
        if (OP == 2) x[i] = __fmul2_rn(x[i], m);
        if (OP == 3) x[i] = __fadd2_rn(x[i], c);
        if (OP == 4) x[i] = __ffma2_rn(x[i], y[i], z[(i + 1) % ILP]);      // three register-pair operands
        if (OP == 5) x[i].x = __fmaf_rn(x[i].x, y[i].x, z[(i + 1) % ILP].x);
        if (OP == 6) {                                                    // the accumulation's shape: acc += a * b
          x[i] = __ffma2_rn(y[i], z[(i + 3) % ILP], x[i]);
        }
      }
  }
  const long long t1 = clock64();
  float2 s = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < ILP; i++) { s.x += x[i].x + y[i].x + z[i].y; s.y += x[i].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP, int ILP>
static double run(int warps_per_smsp) {
  float2* out; long long* cyc;
  const int threads = 128 * warps_per_smsp;        // warp w runs on sub-partition w % 4
  cudaMalloc(&out, sizeof(float2) * threads);
  cudaMalloc(&cyc, sizeof(long long));
  probe<OP, ILP><<<1, threads>>>(out, cyc, make_float2(1.f, 2.f));
  probe<OP, ILP><<<1, threads>>>(out, cyc, make_float2(1.f, 2.f));
  long long h = 0;
  cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  cudaFree(out); cudaFree(cyc);
  return (double)h / ((double)N * ILP * warps_per_smsp);   // cycles per warp-instruction on one sub-partition
}

int main() {
  const char* names[7] = {"FFMA ", "FFMA2", "FMUL2", "FADD2", "FFMA2_rrr", "FFMA_rrr", "FFMA2_acc"};
  printf("{\n");
  for (int op = 0; op < 7; op++) {
    double lat, t[5];
    switch (op) {
      case 0: lat = run<0, 1>(1); t[0] = run<0, 8>(1); t[1] = run<0, 8>(2); t[2] = run<0, 8>(3); t[3] = run<0, 8>(4); t[4] = run<0, 2>(3); break;
      case 1: lat = run<1, 1>(1); t[0] = run<1, 8>(1); t[1] = run<1, 8>(2); t[2] = run<1, 8>(3); t[3] = run<1, 8>(4); t[4] = run<1, 2>(3); break;
      case 2: lat = run<2, 1>(1); t[0] = run<2, 8>(1); t[1] = run<2, 8>(2); t[2] = run<2, 8>(3); t[3] = run<2, 8>(4); t[4] = run<2, 2>(3); break;
      case 4: lat = run<4, 1>(1); t[0] = run<4, 8>(1); t[1] = run<4, 8>(2); t[2] = run<4, 8>(3); t[3] = run<4, 8>(4); t[4] = run<4, 2>(3); break;
      case 5: lat = run<5, 1>(1); t[0] = run<5, 8>(1); t[1] = run<5, 8>(2); t[2] = run<5, 8>(3); t[3] = run<5, 8>(4); t[4] = run<5, 2>(3); break;
      case 6: lat = run<6, 1>(1); t[0] = run<6, 8>(1); t[1] = run<6, 8>(2); t[2] = run<6, 8>(3); t[3] = run<6, 8>(4); t[4] = run<6, 2>(3); break;
      default: lat = run<3, 1>(1); t[0] = run<3, 8>(1); t[1] = run<3, 8>(2); t[2] = run<3, 8>(3); t[3] = run<3, 8>(4); t[4] = run<3, 2>(3); break;
    }
    printf(" \"%s\": {\"latency_cycles\": %.2f, \"cycles_per_warp_instr_ilp8_warps_1_2_3_4\": [%.2f, %.2f, %.2f, %.2f], \"ilp2_3warps\": %.2f}%s\n",
           names[op], lat, t[0], t[1], t[2], t[3], t[4], op < 6 ? "," : "");
  }
  printf("}\n");
  return 0;
}
