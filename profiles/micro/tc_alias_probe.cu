// tc_alias_probe.cu — does a K-major, no-swizzle B descriptor with LEADING BYTE OFFSET 16 work?
// With LBO = 16 the 8 rows of the N = 8 operand alias shifted copies of ONE natural-order vector: row r, piece j = bytes
// [(j + r) * 16, (j + r) * 16 + 16) of x.  Row 0 is x itself, so a single activation row needs no 8-row image in shared memory
// (4 KB instead of 32 KB for K = 4096).  The probe runs one CTA: A = 128 x K int8 rows in the canonical layout, B = natural x,
// K / 32 MMAs accumulated into one TMEM tile, column 0 compared with the CPU dot products (and columns 1-7 with the shifted dots).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ../../fast-llama_b200/csrc -o tc_alias_probe tc_alias_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc_gemm.cuh"
using namespace fl;

constexpr int K = 256, R = 100;

__global__ void __launch_bounds__(128, 1) probe(const int8_t* a_img, const int8_t* x, int* out /* [128][8] */) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    uint8_t* a_s = sm;                                   // [K/16][R] x 16 B
    uint8_t* x_s = sm + (K / 16) * R * 16 + 2048;        // natural x, then 128 B of tail
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (K / 16) * R * 16; i += 128) a_s[i] = (uint8_t)a_img[i];
    for (int i = tid; i < K + 128; i += 128) x_s[i] = i < K ? (uint8_t)x[i] : 0;
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tc_alloc(&tslot, 32);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> tensor core (async proxy) reads
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tslot;
    if (tid == 0) {
        const uint32_t idesc = tc_idesc_i8(128, 8);
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);
        const uint32_t a0 = (smem_u32(a_s) & 0x3ffffu) >> 4, b0 = (smem_u32(x_s) & 0x3ffffu) >> 4;
        for (int k32 = 0; k32 < K / 32; ++k32) {
            const uint32_t a_lo = (a0 + (uint32_t)(k32 * 2) * R) | ((uint32_t)R << 16);
            const uint32_t b_lo = (b0 + (uint32_t)(k32 * 2)) | (1u << 16);          // LBO = 16 bytes: pieces overlap
            if (k32 == 0) tc_mma_i8<false>(tmem, a_lo, b_lo, desc_hi, idesc); else tc_mma_i8<true>(tmem, a_lo, b_lo, desc_hi, idesc);
        }
        tc_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    int d[8];
    tc_ld8(tmem + ((uint32_t)(warp * 32) << 16), d);
    tc_ld_wait();
    for (int i = 0; i < 8; ++i) out[tid * 8 + i] = d[i];
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tc_dealloc(tmem, 32);
}

int main() {
    std::vector<int8_t> w(R * K), x(K + 128, 0), img((K / 16) * R * 16);
    srand(1);
    for (auto& v : w) v = (int8_t)(rand() % 255 - 127);
    for (int i = 0; i < K; ++i) x[i] = (int8_t)(rand() % 255 - 127);
    for (int j = 0; j < K / 16; ++j) for (int r = 0; r < R; ++r) for (int b = 0; b < 16; ++b) img[(j * R + r) * 16 + b] = w[r * K + j * 16 + b];
    int8_t *da, *dx; int* dout;
    cudaMalloc(&da, img.size()); cudaMalloc(&dx, K); cudaMalloc(&dout, 128 * 8 * 4);
    cudaMemcpy(da, img.data(), img.size(), cudaMemcpyHostToDevice); cudaMemcpy(dx, x.data(), K, cudaMemcpyHostToDevice);
    const size_t smem = (K / 16) * R * 16 + 2048 + K + 256;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<<<1, 128, smem>>>(da, dx, dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<int> out(128 * 8);
    cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
    int bad0 = 0, badn = 0;
    for (int r = 0; r < R; ++r) for (int c = 0; c < 8; ++c) {
        long ref = 0;
        for (int k = 0; k < K; ++k) ref += (int)w[r * K + k] * (int)x[k + 16 * c];       // column c = x shifted by 16 c bytes
        if ((int)ref != out[r * 8 + c]) { if (c == 0) ++bad0; else ++badn; if (bad0 + badn < 6) printf("row %d col %d: got %d want %ld\n", r, c, out[r * 8 + c], ref); }
    }
    printf("LBO=16 alias probe: column 0 mismatches %d / %d, shifted columns mismatches %d / %d  -> %s\n", bad0, R, badn, R * 7, bad0 == 0 ? "OK" : "FAIL");
    return bad0 != 0;
}
