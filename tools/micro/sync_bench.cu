// Micro-benchmark (development aid, not part of the product): what does a phase boundary cost on this GPU?
//  (a) dependent kernel nodes in a CUDA graph: us per node for an (almost) empty kernel, with and without PDL;
//  (b) a hand-written grid barrier (atomic arrive + spin on an L2 word) inside one co-resident kernel: us per barrier,
//      for 1, 2, 3, 4 CTAs of 256 threads per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/sync_bench tools/micro/sync_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include <algorithm>

__global__ void tiny_kernel(unsigned* p) { if (threadIdx.x == 0 && blockIdx.x == 0) p[0] += 1u; }

__global__ void tiny_kernel_pdl(unsigned* p) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x == 0 && blockIdx.x == 0) p[0] += 1u;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& phase_target, unsigned nctas) {
    __syncthreads();
    if (threadIdx.x == 0) {
        phase_target += nctas;
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while (v < phase_target);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) barrier_kernel(unsigned* counter, unsigned* out, int nbar) {
    unsigned target = 0;
    for (int i = 0; i < nbar; ++i) grid_barrier(counter, target, gridDim.x);
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = target;
}

static float time_graph(cudaGraphExec_t g, cudaStream_t s, int reps) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    std::vector<float> ts;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(a, s); cudaGraphLaunch(g, s); cudaEventRecord(b, s); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); ts.push_back(ms);
    }
    std::sort(ts.begin(), ts.end());
    return ts[ts.size() / 2];
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    unsigned* d; cudaMalloc(&d, 64); cudaMemset(d, 0, 64);
    cudaStream_t s; cudaStreamCreate(&s);
    const int N = 64;
    for (int grid_mul = 0; grid_mul <= 2; grid_mul += 2) {
        const int grid = grid_mul == 0 ? 1 : sms * grid_mul;
        for (int pdl = 0; pdl < 2; ++pdl) {
            cudaGraph_t graph; cudaGraphExec_t exec;
            cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
            for (int i = 0; i < N; ++i) {
                cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.stream = s;
                cudaLaunchAttribute attr[1]; attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                attr[0].val.programmaticStreamSerializationAllowed = 1; cfg.attrs = attr; cfg.numAttrs = pdl;
                if (pdl) cudaLaunchKernelEx(&cfg, tiny_kernel_pdl, d); else cudaLaunchKernelEx(&cfg, tiny_kernel, d);
            }
            cudaStreamEndCapture(s, &graph); cudaGraphInstantiate(&exec, graph, 0);
            time_graph(exec, s, 3);
            printf("graph of %d dependent tiny kernels, grid %d, pdl %d: %.2f us per node\n", N, grid, pdl, time_graph(exec, s, 9) * 1e3f / N);
            cudaGraphExecDestroy(exec); cudaGraphDestroy(graph);
        }
    }
    for (int per_sm = 1; per_sm <= 4; ++per_sm) {
        for (int nbar : {0, 32}) {
            cudaMemsetAsync(d, 0, 64, s);
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            std::vector<float> ts;
            for (int r = 0; r < 9; ++r) {
                cudaMemsetAsync(d, 0, 64, s);
                void* args[] = {(void*)&d, (void*)&d, (void*)&nbar};
                unsigned* out = d + 8; args[1] = &out;
                cudaEventRecord(a, s);
                cudaLaunchCooperativeKernel((void*)barrier_kernel, dim3(sms * per_sm), dim3(256), args, 0, s);
                cudaEventRecord(b, s); cudaEventSynchronize(b);
                float ms; cudaEventElapsedTime(&ms, a, b); ts.push_back(ms);
            }
            std::sort(ts.begin(), ts.end());
            printf("cooperative kernel, %d CTAs/SM, %d grid barriers: %.2f us total%s\n", per_sm, nbar, ts[4] * 1e3f,
                   nbar ? "" : " (launch only)");
            if (nbar) printf("   -> %.2f us per barrier\n", (ts[4] * 1e3f) / nbar);
        }
    }
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
