// tools/cub_sort_yardstick.cu -- YARDSTICK ONLY: times CUB's DeviceRadixSort::SortPairs on the
// same key shape the collision step sorts (u32 cell key, u32 collider index, `bits` significant key bits), so that
// gsc_sort.cuh can be judged against a tuned library sort on the same box.  Not linked into the product library.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/cub_sort_yardstick tools/cub_sort_yardstick.cu
//   tools/cub_sort_yardstick [n=16777216] [bits=27] [reps=20]
#include <cub/cub.cuh>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

int main(int argc, char **argv)
{
    const size_t n = argc > 1 ? strtoull(argv[1], nullptr, 10) : (1ull << 24);
    const int bits = argc > 2 ? atoi(argv[2]) : 27;
    const int reps = argc > 3 ? atoi(argv[3]) : 20;
    std::vector<uint32_t> hk(n);
    uint64_t s = 0x9E3779B97F4A7C15ull;
    for (size_t i = 0; i < n; ++i) {  // SplitMix64
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        hk[i] = (uint32_t)((z ^ (z >> 31)) & ((bits >= 32 ? 0ull : (1ull << bits)) - 1ull));
    }
    uint32_t *k[2], *v[2], *src;
    for (int i = 0; i < 2; ++i) { CK(cudaMalloc(&k[i], n * 4)); CK(cudaMalloc(&v[i], n * 4)); }
    CK(cudaMalloc(&src, n * 4));
    CK(cudaMemcpy(src, hk.data(), n * 4, cudaMemcpyHostToDevice));
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    cub::DoubleBuffer<uint32_t> dk(k[0], k[1]), dv(v[0], v[1]);
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, (int)n, 0, bits));
    CK(cudaMalloc(&tmp, tmp_bytes));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f, sum = 0;
    for (int r = 0; r < reps + 3; ++r) {
        cub::DoubleBuffer<uint32_t> a(k[0], k[1]), b(v[0], v[1]);
        CK(cudaMemcpy(k[0], src, n * 4, cudaMemcpyDeviceToDevice));  // (also evicts nothing: 64 MB < L2 -- the step's sort input is L2-warm too)
        CK(cudaMemset(v[0], 0, n * 4));
        CK(cudaEventRecord(e0));
        CK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, a, b, (int)n, 0, bits));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r >= 3) { best = ms < best ? ms : best; sum += ms; }
    }
    const int passes = (bits + 7) / 8;
    printf("{\"what\": \"cub::DeviceRadixSort::SortPairs (yardstick)\", \"cub_version\": %d, \"n\": %zu, \"key_bits\": %d, \"best_ms\": %.4f, "
           "\"mean_ms\": %.4f, \"gkeys_per_s\": %.2f, \"alg_bytes\": %zu, \"alg_gbs\": %.1f}\n",
           CUB_VERSION, n, bits, best, sum / reps, n / best / 1e6, n * 16 * (size_t)passes, n * 16.0 * passes / best / 1e6);
    return 0;
}
