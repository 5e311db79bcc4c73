// Micro-benchmark: cycles per row of one dependent float chain per lane fed from shared memory
// (the inner loop of the STRICT reduction of REFERENCE mode).  Variants: scalar loads interleaved with
// the adds, 128-bit loads, and registers only (the FADD latency floor).
#include <cstdio>
#include <cuda_runtime.h>
constexpr int kStride = 1028;  // floats per column: 16-byte aligned columns
__global__ void chain(const float *in, float *out, long long *cyc, int n, int lanes, int variant)
{
    __shared__ __align__(16) float s[9 * kStride];
    for (int i = threadIdx.x; i < 9 * kStride; i += blockDim.x) s[i] = in[i % 1024];
    __syncthreads();
    if (threadIdx.x >= 32) return;
    float acc = 0.f;
    long long t0 = clock64();
    if (threadIdx.x < lanes) {
        const float *col = s + threadIdx.x * kStride;
        for (int rep = 0; rep < n; ++rep) {
            if (variant == 0) {  // scalar loads, two register sets
                float a[16], b[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) a[u] = col[u];
                for (int j = 0; j < 1024; j += 32) {
#pragma unroll
                    for (int u = 0; u < 16; ++u) b[u] = col[j + 16 + u];
#pragma unroll
                    for (int u = 0; u < 16; ++u) acc = __fadd_rn(acc, a[u]);
                    if (j + 32 < 1024) {
#pragma unroll
                        for (int u = 0; u < 16; ++u) a[u] = col[j + 32 + u];
                    }
#pragma unroll
                    for (int u = 0; u < 16; ++u) acc = __fadd_rn(acc, b[u]);
                }
            } else if (variant == 1) {  // 128-bit loads
                const float4 *c4 = reinterpret_cast<const float4 *>(col);
                float4 a[4], b[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) a[u] = c4[u];
                for (int j = 0; j < 256; j += 8) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) b[u] = c4[j + 4 + u];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        acc = __fadd_rn(acc, a[u].x); acc = __fadd_rn(acc, a[u].y);
                        acc = __fadd_rn(acc, a[u].z); acc = __fadd_rn(acc, a[u].w);
                    }
                    if (j + 8 < 256) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) a[u] = c4[j + 8 + u];
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        acc = __fadd_rn(acc, b[u].x); acc = __fadd_rn(acc, b[u].y);
                        acc = __fadd_rn(acc, b[u].z); acc = __fadd_rn(acc, b[u].w);
                    }
                }
            } else if (variant == 3) {  // 128-bit loads, three register sets, 48 rows per trip
                const float4 *c4 = reinterpret_cast<const float4 *>(col);
                float4 a[4], b[4], c[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { a[u] = c4[u]; b[u] = c4[4 + u]; }
                int j = 0;
                for (; j + 12 <= 256; j += 12) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) c[u] = c4[min(j + 8 + u, 256)];
#pragma unroll
                    for (int u = 0; u < 4; ++u) { acc = __fadd_rn(acc, a[u].x); acc = __fadd_rn(acc, a[u].y); acc = __fadd_rn(acc, a[u].z); acc = __fadd_rn(acc, a[u].w); }
#pragma unroll
                    for (int u = 0; u < 4; ++u) a[u] = c4[min(j + 12 + u, 256)];
#pragma unroll
                    for (int u = 0; u < 4; ++u) { acc = __fadd_rn(acc, b[u].x); acc = __fadd_rn(acc, b[u].y); acc = __fadd_rn(acc, b[u].z); acc = __fadd_rn(acc, b[u].w); }
#pragma unroll
                    for (int u = 0; u < 4; ++u) b[u] = c4[min(j + 16 + u, 256)];
#pragma unroll
                    for (int u = 0; u < 4; ++u) { acc = __fadd_rn(acc, c[u].x); acc = __fadd_rn(acc, c[u].y); acc = __fadd_rn(acc, c[u].z); acc = __fadd_rn(acc, c[u].w); }
                }
                for (; j < 256; ++j) { const float4 v = c4[j]; acc = __fadd_rn(acc, v.x); acc = __fadd_rn(acc, v.y); acc = __fadd_rn(acc, v.z); acc = __fadd_rn(acc, v.w); }
            } else if (variant == 4) {  // 128-bit loads, fully unrolled 256 rows per trip, loads 32 rows ahead
                const float4 *c4 = reinterpret_cast<const float4 *>(col);
                for (int j0 = 0; j0 < 256; j0 += 64) {
                    float4 r[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) r[u] = c4[j0 + u];
#pragma unroll
                    for (int u = 0; u < 64; ++u) {
                        const float4 v = r[u & 7];
                        if (u + 8 < 64) r[u & 7] = c4[j0 + u + 8];
                        acc = __fadd_rn(acc, v.x); acc = __fadd_rn(acc, v.y); acc = __fadd_rn(acc, v.z); acc = __fadd_rn(acc, v.w);
                    }
                }
            } else {  // registers only
                float a[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) a[u] = col[u];
                for (int j = 0; j < 1024; j += 16) {
#pragma unroll
                    for (int u = 0; u < 16; ++u) acc = __fadd_rn(acc, a[u]);
                }
            }
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = acc;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main()
{
    float *in, *out; long long *cyc;
    cudaMalloc(&in, 4096); cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
    cudaMemset(in, 0, 4096);
    for (int variant = 0; variant < 6; ++variant)
        for (int lanes : {1, 9}) {
            chain<<<1, 64>>>(in, out, cyc, 64, lanes, variant);
            long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            printf("variant %d lanes %d: %.2f cycles per row\n", variant, lanes, (double)c / (64.0 * 1024));
        }
    return 0;
}
