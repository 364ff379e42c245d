// Analysis tool: latency probes on one warp — dependent chains of the building blocks of the quad RHS (cycles per
// link, including ~15 cycles of loop overhead).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/lat scripts/latency_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../simplyp_b200/csrc/simplyp_quad.cuh"
using namespace simplyp;
#define N 512
template <int WHICH>
__global__ void probe(double* out, long long* cyc, double a, double b) {
  __shared__ double tab[64];
  tab[threadIdx.x] = kExp2Tab[threadIdx.x]; tab[threadIdx.x + 32] = kExp2Tab[threadIdx.x + 32];
  __syncthreads();
  QuadDev q; q.ql = threadIdx.x & 3; q.tab = tab;
  QuadCoef<QuadDev> c;
  Hot h; memset(&h, 0, sizeof(h));
  h.fc = 290; h.inv_fcd = 1 / 2.9; h.mu = 0.0159; h.inv_TsA = 0.5; h.inv_TsS = 0.1; h.inv_Tg = 1 / 65.; h.Qg_min = 0.4; h.inv_Qgd = 1 / 0.004;
  h.Pin = 3; h.aE = 1; h.fA = .5; h.fS = .5; h.beta = .7; h.qin0 = 0.1; h.kQ = 7; h.bQ = 0.42; h.kM = 2; h.cR = 4.3; h.cM = 300; h.tA = .3; h.tS = .1; h.tG = 1; h.t0 = .5; h.cP = 2;
  quad_static_coef(q, h, c); quad_daily_coef(q, h, c);
  double x = a + threadIdx.x * 1e-3, y = b;
  double yA = q.pick(291.0, 290.5, 80.0, 0.3), yB = q.pick(10.0, 0.1, 0.1, 0.35);
  double yA2 = q.pick(292.0, 291.5, 70.0, 0.5), yB2 = q.pick(12.0, 0.2, 0.2, 0.30);
  QuadCoef<QuadDev> c2;
  h.Pin = 5; h.aE = 0.5; h.qin0 = 0.2;
  quad_static_coef(q, h, c2); quad_daily_coef(q, h, c2);
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) {
    if (WHICH == 0) x = fma(x, a, b);
    if (WHICH == 1) x = q.bcast(x, 3) + 1e-9;
    if (WHICH == 2) x = q.exp(x * 1e-3);
    if (WHICH == 3) x = qrcp(x) + 1.0;
    if (WHICH == 4) x = qgate(x * 0.5) + 0.1;
    if (WHICH == 5) { double dA, dB, da, e; quad_rhs(q, c, yA, yB, dA, dB, da, e); yA = fma(1e-9, dA, yA); yB = fma(1e-9, dB, yB); x = yA; }
    if (WHICH == 10) {   // two independent members per quad, interleaved in one instruction stream
      double dA, dB, da, e, dA2, dB2, da2, e2;
      quad_rhs(q, c, yA, yB, dA, dB, da, e);
      quad_rhs(q, c2, yA2, yB2, dA2, dB2, da2, e2);
      yA = fma(1e-9, dA, yA); yB = fma(1e-9, dB, yB); yA2 = fma(1e-9, dA2, yA2); yB2 = fma(1e-9, dB2, yB2); x = yA + yA2;
    }
    if (WHICH == 6) x = x * a;
    if (WHICH == 7) x = x + b;
    if (WHICH == 8) x = step_factor_sq(x + 1.0) ;
    if (WHICH == 9) { double s = x; s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2); x = s * 0.25; }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = x + y;
}
int main() {
  double* d; long long* c; cudaMalloc(&d, 256 * 8); cudaMalloc(&c, 8);
  const char* names[] = {"DFMA", "bcast64 (2 SHFL) + DADD", "table exp (+DMUL)", "rcp (MUFU+2 Newton) + DADD", "gate (+DMUL,DADD)", "quad_rhs + 2 DFMA", "DMUL", "DADD", "step_factor_sq (+DADD)", "quad sum (2 rounds)", "2 interleaved quad_rhs + 4 DFMA"};
  long long hc;
#define RUN(W) for (int r = 0; r < 2; ++r) { probe<W><<<1, 32>>>(d, c, 0.999999, 1e-9); cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost); } printf("%-32s %7.1f cycles per link\n", names[W], (double)hc / N);
  RUN(0) RUN(6) RUN(7) RUN(1) RUN(9) RUN(2) RUN(3) RUN(4) RUN(8) RUN(5) RUN(10)
  return 0;
}
