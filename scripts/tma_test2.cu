#include <cstdio>
#include <vector>
#include "../gknextengine_b200/csrc/gk_tma.cuh"
constexpr int TW = 36, TH = 12, MW = 40, MH = 16;
struct Maps { CUtensorMap a, b, c, ids, m; };
struct Args { int w, h; const void* p[12]; unsigned x, y; int z; };
__global__ void __launch_bounds__(256) k(const __grid_constant__ Maps M, Args A, unsigned long long* out, int mask) {
    __shared__ alignas(128) uint2 ra[TH][TW];
    __shared__ alignas(128) uint2 rb[TH][TW];
    __shared__ alignas(128) uint2 rc[TH][TW];
    __shared__ alignas(128) uint32_t sid[TH][TW];
    __shared__ alignas(128) float2 rm[MH][MW];
    __shared__ uint64_t bar;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int bx0 = blockIdx.x * 32, by0 = blockIdx.y * 8;
    if (tid == 0) gktma::mbar_init(&bar, 1);
    __syncthreads();
    if (tid == 0) {
        uint32_t bytes = 0;
        if (mask & 1) bytes += sizeof(ra);
        if (mask & 2) bytes += sizeof(rb);
        if (mask & 4) bytes += sizeof(rc);
        if (mask & 8) bytes += sizeof(sid);
        if (mask & 16) bytes += sizeof(rm);
        gktma::mbar_expect_tx(&bar, bytes);
        if (mask & 1) gktma::tma_load_2d(ra, &M.a, bx0 - 2, by0 - 2, &bar);
        if (mask & 2) gktma::tma_load_2d(rb, &M.b, bx0 - 2, by0 - 2, &bar);
        if (mask & 4) gktma::tma_load_2d(rc, &M.c, bx0 - 2, by0 - 2, &bar);
        if (mask & 8) gktma::tma_load_2d(sid, &M.ids, bx0 - 2, by0 - 2, &bar);
        if (mask & 16) gktma::tma_load_2d(rm, &M.m, bx0 - 4, by0 - 4, &bar);
    }
    gktma::mbar_wait(&bar, 0);
    if (blockIdx.x == 1 && blockIdx.y == 1 && tid < 64) out[tid] = ((unsigned long long)sid[2][2 + (tid & 31)] << 32) | ra[2][2 + (tid & 31)].x;
}
int main() {
    const int W = 1920, H = 1080;
    void *a, *b, *c, *ids, *m; unsigned long long* o;
    cudaMalloc(&a, (size_t)W * H * 8); cudaMalloc(&b, (size_t)W * H * 8); cudaMalloc(&c, (size_t)W * H * 8); cudaMalloc(&ids, (size_t)W * H * 4); cudaMalloc(&m, (size_t)W * H * 8);
    cudaMalloc(&o, 64 * 8);
    cudaMemset(a, 1, (size_t)W * H * 8); cudaMemset(ids, 2, (size_t)W * H * 4);
    Maps M;
    bool ok = gktma::encode_2d(&M.a, a, CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, W, H, TW, TH) && gktma::encode_2d(&M.b, b, CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, W, H, TW, TH) &&
              gktma::encode_2d(&M.c, c, CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, W, H, TW, TH) && gktma::encode_2d(&M.ids, ids, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, W, H, TW, TH) &&
              gktma::encode_2d(&M.m, m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, W, H, MW, MH);
    printf("encode %d\n", ok);
    Args A{};
    for (int mask : {1, 2, 4, 8, 16, 3, 7, 15, 31}) {
        k<<<dim3(60, 135), dim3(32, 8)>>>(M, A, o, mask);
        cudaError_t e = cudaDeviceSynchronize();
        printf("mask %2d sync: %s\n", mask, cudaGetErrorString(e));
        if (e != cudaSuccess) break;
    }
    unsigned long long r[2]; cudaMemcpy(r, o, 16, cudaMemcpyDeviceToHost);
    printf("%llx\n", r[0]);
    return 0;
}
