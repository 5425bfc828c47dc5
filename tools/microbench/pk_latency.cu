// Latency / issue rate of the packed fp32x2 forms next to the scalar ones on sm_100a (one B200).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pk_latency pk_latency.cu && ./pk_latency
// lat: one warp, one dependent chain -> cycles per instruction = result latency.
// rate: 4 warps (one per SM sub-partition) x 8 independent chains -> cycles per instruction per sub-partition.
#include <cstdio>
#include <cuda_runtime.h>

#define N 4096
template <int OP, int ILP>
__global__ void k(float2 *out, long long *cyc, float a, float b) {
    float2 x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = make_float2(threadIdx.x * 0.001f + i, 1.0f + i);
    const float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b * 0.9999f);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < N / 16; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (OP == 0) x[i].x = fmaf(x[i].x, a2.x, b2.x);                    // FFMA
                else if (OP == 1) x[i] = __ffma2_rn(x[i], a2, b2);                 // FFMA2
                else if (OP == 2) x[i].x = x[i].x + b2.x;                          // FADD
                else if (OP == 3) x[i] = __fadd2_rn(x[i], b2);                     // FADD2
                else if (OP == 4) x[i].x = x[i].x * a2.x;                          // FMUL
                else if (OP == 5) x[i] = __fmul2_rn(x[i], a2);                     // FMUL2
                else if (OP == 6) x[i].x = fmaxf(x[i].x * 1.0f, b2.x), x[i].x = fminf(x[i].x, a2.y + i);   // 2 x FMNMX (ALU pipe)
            }
        }
    }
    const long long t1 = clock64();
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < ILP; ++i) { s.x += x[i].x; s.y += x[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP, int ILP>
static double run(int threads, int per_iter) {
    float2 *out; long long *cyc, h = 0;
    cudaMalloc(&out, sizeof(float2) * 1024); cudaMalloc(&cyc, sizeof(long long) * 8);
    for (int rep = 0; rep < 3; ++rep) k<OP, ILP><<<1, threads>>>(out, cyc, 0.999f, 0.001f);
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(out); cudaFree(cyc);
    return (double)h / ((double)N * ILP * per_iter);
}

int main() {
    const char *names[] = {"FFMA", "FFMA2", "FADD", "FADD2", "FMUL", "FMUL2", "FMNMX"};
    double lat[7], rate[7];
    lat[0] = run<0, 1>(32, 1); lat[1] = run<1, 1>(32, 1); lat[2] = run<2, 1>(32, 1); lat[3] = run<3, 1>(32, 1);
    lat[4] = run<4, 1>(32, 1); lat[5] = run<5, 1>(32, 1); lat[6] = run<6, 1>(32, 2);
    rate[0] = run<0, 8>(128, 1); rate[1] = run<1, 8>(128, 1); rate[2] = run<2, 8>(128, 1); rate[3] = run<3, 8>(128, 1);
    rate[4] = run<4, 8>(128, 1); rate[5] = run<5, 8>(128, 1); rate[6] = run<6, 8>(128, 2);
    printf("%-6s %22s %40s\n", "op", "latency [cycles]", "issue interval, 1 warp / sub-partition, ILP 8");
    for (int i = 0; i < 7; ++i) printf("%-6s %22.2f %40.2f\n", names[i], lat[i], rate[i]);
    return 0;
}
