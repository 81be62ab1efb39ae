// locbw.cu — design microbenchmark (not part of the product library).
//
// Question: if the dictionary bucket of a k-mer were chosen by its MINIMIZER, neighbouring k-mers of a read would share
// a bucket, and the lanes of a warp that hold them would ask for the same sectors — which the load unit merges into one
// request. How many "k-mers"/s does the memory system then deliver? Pattern per k-mer (a lane): one 8-byte directory
// entry at a random bucket, then — dependent on it — one 32-byte sector of the bucket's rows (a random 128-byte line;
// the lanes of a group read different sectors of that line). Groups of G consecutive lanes share the bucket; G = 1 is
// today's pattern with a directory in front (two dependent requests per k-mer), `flat` is today's dictionary itself
// (one random sector per k-mer, no directory).
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o locbw locbw.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
__device__ __forceinline__ uint64_t ld32(const uint8_t *p) {
    uint64_t a, b, c, d;
    asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    return a ^ b ^ c ^ d;
}
__device__ __forceinline__ uint64_t ld8(const uint8_t *p) {
    uint64_t a;
    asm volatile("ld.global.nc.L1::no_allocate.L2::64B.u64 %0, [%1];" : "=l"(a) : "l"(p));
    return a;
}

// today's dictionary: one random sector per k-mer
__global__ void flat(const uint8_t *rows, uint64_t nsectors, int hops, uint64_t *sink) {
    uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t st = mix(tid + 1);
    for (int h = 0; h < hops; ++h) st = mix(st + ld32(rows + (st % nsectors) * 32));
    if (st == 0x1234567) sink[0] = st;
}

// directory entry -> rows sector; G consecutive lanes share the bucket. SECT = sectors of the line a group spreads over.
template <int G>
__global__ void grouped(const uint8_t *dir, uint64_t ndir, const uint8_t *rows, uint64_t nlines, int hops, int sect, uint64_t *sink) {
    uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t st = mix(tid / G + 1);  // the same chain for the G lanes of a group
    const uint32_t mysect = (threadIdx.x % G) % sect;
    uint64_t acc = 0;
    for (int h = 0; h < hops; ++h) {
        const uint64_t d = ld8(dir + (st % ndir) * 8);
        const uint64_t line = mix(st ^ d) % nlines;  // dependent on the directory entry (all zeroes)
        const uint64_t v = ld32(rows + line * 128 + mysect * 32);
        acc += v;
        st = mix(st + v);  // rows are all zeroes: the lanes of a group stay on the same chain, and the next hop waits for this one
    }
    if (st + acc == 0x1234567) sink[0] = st;
}

template <typename F>
float time_it(F launch) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(1);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        launch(0);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main(int argc, char **argv) {
    CK(cudaSetDevice(0));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int nsm = p.multiProcessorCount;
    const uint64_t dir_bytes = (argc > 1 ? strtoull(argv[1], 0, 10) : 8192ull) << 20;    // 2^30 entries of 8 B
    const uint64_t rows_bytes = (argc > 2 ? strtoull(argv[2], 0, 10) : 24576ull) << 20;  // 3.1 G rows of 8 B
    const int hops = argc > 3 ? atoi(argv[3]) : 64;
    uint8_t *dir, *rows; uint64_t *sink;
    CK(cudaMalloc(&dir, dir_bytes)); CK(cudaMalloc(&rows, rows_bytes)); CK(cudaMalloc(&sink, 8));
    CK(cudaMemset(dir, 0, dir_bytes)); CK(cudaMemset(rows, 0, rows_bytes));
    printf("# %s, %d SMs, directory %.1f GB, rows %.1f GB, %d hops per lane\n", p.name, nsm, dir_bytes / 1e9, rows_bytes / 1e9, hops);
    for (int tps : {1024, 1280, 2048}) {
        const int bs = 256, grid = nsm * (tps / bs);
        const double kmers = (double)grid * bs * hops;
        float ms = time_it([&](int warm) { flat<<<grid, bs>>>(rows, rows_bytes / 32, warm ? hops / 4 : hops, sink); });
        printf("{\"pattern\": \"flat (one random sector per k-mer)\", \"threads_per_sm\": %d, \"ms\": %.3f, \"gkmers_s\": %.2f}\n", tps, ms, kmers / ms / 1e6);
        for (int sect : {1, 4}) {
#define RUN(G) { float t = time_it([&](int warm) { grouped<G><<<grid, bs>>>(dir, dir_bytes / 8, rows, rows_bytes / 128, warm ? hops / 4 : hops, sect, sink); }); \
                 printf("{\"pattern\": \"directory + rows line\", \"group\": %d, \"sectors_of_line\": %d, \"threads_per_sm\": %d, \"ms\": %.3f, \"gkmers_s\": %.2f, \"grequests_s\": %.2f}\n", \
                        G, sect, tps, t, kmers / t / 1e6, 2.0 * kmers / G / t / 1e6); fflush(stdout); }
            RUN(1) RUN(2) RUN(4) RUN(8) RUN(16) RUN(32)
#undef RUN
        }
    }
    return 0;
}
