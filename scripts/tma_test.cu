// Stand-alone check of the TMA staging helpers (gk_tma.cuh): loads tiles with halo from a W x H image of 8-byte pixels.
#include <cstdio>
#include <vector>
#include "../gknextengine_b200/csrc/gk_tma.cuh"
constexpr int TW = 36, TH = 12;
__global__ void k(const __grid_constant__ CUtensorMap map, unsigned long long* out, int x0, int y0) {
    __shared__ alignas(128) unsigned long long tile[TH][TW];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) gktma::mbar_init(&bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        gktma::mbar_expect_tx(&bar, sizeof(tile));
        gktma::tma_load_2d(tile, &map, x0, y0, &bar);
    }
    gktma::mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < TW * TH; i += blockDim.x) out[i] = tile[i / TW][i % TW];
}
int main() {
    const int W = 64, H = 32;
    std::vector<unsigned long long> h(W * H);
    for (int i = 0; i < W * H; ++i) h[i] = 1000 + i;
    unsigned long long *d, *o;
    cudaMalloc(&d, W * H * 8); cudaMalloc(&o, TW * TH * 8);
    cudaMemcpy(d, h.data(), W * H * 8, cudaMemcpyHostToDevice);
    CUtensorMap map;
    bool ok = gktma::encode_2d(&map, d, CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, W, H, TW, TH);
    printf("encode %d\n", ok);
    k<<<1, 128>>>(map, o, -2, -2);
    cudaError_t e = cudaDeviceSynchronize();
    printf("sync: %s\n", cudaGetErrorString(e));
    std::vector<unsigned long long> r(TW * TH);
    cudaMemcpy(r.data(), o, TW * TH * 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int y = 0; y < TH; ++y) for (int x = 0; x < TW; ++x) {
        int gx = x - 2, gy = y - 2;
        unsigned long long expect = (gx >= 0 && gy >= 0 && gx < W && gy < H) ? 1000 + gy * W + gx : 0;
        if (r[y * TW + x] != expect) ++bad;
    }
    printf("mismatches %d (r[0]=%llu r[2*TW+2]=%llu)\n", bad, r[0], r[2 * TW + 2]);
    return 0;
}
