// Probe: do kernel nodes captured from streams of different priority inherit that priority?
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
__global__ void k_a(int *p) { if (p) p[0] = 1; }
__global__ void k_b(int *p) { if (p) p[1] = 2; }
int main()
{
    int lo, hi;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    printf("priority range lo %d hi %d\n", lo, hi);
    cudaStream_t s0, s1;
    cudaStreamCreateWithPriority(&s0, cudaStreamNonBlocking, lo);
    cudaStreamCreateWithPriority(&s1, cudaStreamNonBlocking, hi);
    cudaEvent_t f, j;
    cudaEventCreateWithFlags(&f, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&j, cudaEventDisableTiming);
    int *d;
    cudaMalloc(&d, 8);
    cudaGraph_t g;
    cudaStreamBeginCapture(s0, cudaStreamCaptureModeThreadLocal);
    cudaEventRecord(f, s0);
    cudaStreamWaitEvent(s1, f, 0);
    k_a<<<1, 32, 0, s0>>>(d);
    k_b<<<1, 32, 0, s1>>>(d);
    cudaEventRecord(j, s1);
    cudaStreamWaitEvent(s0, j, 0);
    cudaStreamEndCapture(s0, &g);
    size_t n = 0;
    cudaGraphGetNodes(g, nullptr, &n);
    std::vector<cudaGraphNode_t> nodes(n);
    cudaGraphGetNodes(g, nodes.data(), &n);
    for (size_t i = 0; i < n; ++i) {
        cudaGraphNodeType t;
        cudaGraphNodeGetType(nodes[i], &t);
        if (t != cudaGraphNodeTypeKernel) continue;
        cudaKernelNodeParams kp;
        cudaGraphKernelNodeGetParams(nodes[i], &kp);
        cudaKernelNodeAttrValue v;
        cudaError_t e = cudaGraphKernelNodeGetAttribute(nodes[i], cudaKernelNodeAttributePriority, &v);
        printf("node %zu func %s priority %d (%s)\n", i, kp.func == (void *)k_a ? "k_a(main,lo)" : "k_b(side,hi)", v.priority, cudaGetErrorString(e));
    }
    return 0;
}
