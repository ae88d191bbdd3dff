// Issue rate of the warp-level (legacy) MMAs on sm_100a: clocks per instruction per SM sub-partition, measured with
// 1..8 warps per sub-partition and 4 independent accumulator chains per warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int KIND>
__global__ void rate_kernel(float *out, long long *clk, int iters) {
    float acc[4][4] = {};
    uint32_t a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(acc[c][0]), "+f"(acc[c][1]), "+f"(acc[c][2]), "+f"(acc[c][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (KIND == 1)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(acc[c][0]), "+f"(acc[c][1]), "+f"(acc[c][2]), "+f"(acc[c][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+f"(acc[c][0]), "+f"(acc[c][1]), "+f"(acc[c][2]), "+f"(acc[c][3])
                             : "r"(a0), "r"(a1), "r"(b0));
        }
    }
    const long long t1 = clock64();
    float s = 0;
    for (int c = 0; c < 4; ++c) for (int k = 0; k < 4; ++k) s += acc[c][k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int KIND>
void run(const char *name, long macs) {
    float *out; long long *clk;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 148 * 8);
    const int iters = 2000;
    for (int warps = 4; warps <= 32; warps *= 2) {
        rate_kernel<KIND><<<148, warps * 32>>>(out, clk, iters);
        rate_kernel<KIND><<<148, warps * 32>>>(out, clk, iters);
        long long h[148];
        cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
        double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
        const double per_smsp = (double)iters * 4 * (warps / 4.0);      // instructions per sub-partition
        printf("%-28s warps/SM %2d: %.2f clk per MMA per sub-partition, %.0f MAC/clk/SM\n", name, warps, c / per_smsp,
               macs * per_smsp * 4 / c);
    }
}

int main() {
    run<0>("m16n8k8 tf32", 16 * 8 * 8);
    run<2>("m16n8k4 tf32", 16 * 8 * 4);
    run<1>("m16n8k16 bf16", 16 * 8 * 16);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
