// probe: cost of a WHILE conditional graph node per loop iteration against eager launches with a sync per round
#include <cuda_runtime.h>
#include <cstdio>
#include <chrono>
#include <vector>
#include <algorithm>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("ERR %s line %d: %s\n", #x, __LINE__, cudaGetErrorString(e_)); return 1; } } while (0)
__global__ void work(double* p, int n) {
  double a = p[threadIdx.x];
  for (int i = 0; i < n; ++i) a = a * 1.0000001 + 1e-9;
  p[threadIdx.x] = a;
}
__global__ void cond(int* counter, int rounds, cudaGraphConditionalHandle h) {
  int c = ++(*counter);
  cudaGraphSetConditional(h, c < rounds ? 1 : 0);
}
__global__ void reset(int* counter) { *counter = 0; }
int main() {
  double* d; int* ctr; int* hflag;
  CK(cudaMalloc(&d, 1024 * 8)); CK(cudaMemset(d, 0, 1024 * 8));
  CK(cudaMalloc(&ctr, 4)); CK(cudaMallocHost(&hflag, 4));
  cudaStream_t st; CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  const int work_n = 4000;  // ~ tens of microseconds of dependent fp64 FMAs
  for (int rounds = 1; rounds <= 3; ++rounds) {
    cudaGraph_t g; CK(cudaGraphCreate(&g, 0));
    CK(cudaStreamBeginCaptureToGraph(st, g, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
    reset<<<1, 1, 0, st>>>(ctr);
    work<<<1, 64, 0, st>>>(d, 100);
    CK(cudaStreamEndCapture(st, &g));
    size_t nl = 0; CK(cudaGraphGetNodes(g, nullptr, &nl));
    std::vector<cudaGraphNode_t> nodes(nl); CK(cudaGraphGetNodes(g, nodes.data(), &nl));
    // last captured node = leaf
    cudaGraphNode_t leaf = nodes[nl - 1];
    size_t nleaf = 0; CK(cudaGraphGetNodes(g, nullptr, &nleaf));
    cudaGraphConditionalHandle h;
    CK(cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault));
    cudaGraphNodeParams p = {};
    p.type = cudaGraphNodeTypeConditional;
    p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
    cudaGraphNode_t cnode;
    // find the real leaf: node without outgoing edges
    for (auto n : nodes) { size_t nd = 0; cudaGraphNodeGetDependentNodes(n, nullptr, &nd); if (nd == 0) leaf = n; }
    CK(cudaGraphAddNode(&cnode, g, &leaf, 1, &p));
    cudaGraph_t body = p.conditional.phGraph_out[0];
    CK(cudaStreamBeginCaptureToGraph(st, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
    for (int k = 0; k < 5; ++k) work<<<1, 64, 0, st>>>(d, work_n);
    cond<<<1, 1, 0, st>>>(ctr, rounds, h);
    CK(cudaStreamEndCapture(st, &body));
    CK(cudaStreamBeginCaptureToGraph(st, g, &cnode, nullptr, 1, cudaStreamCaptureModeRelaxed));
    work<<<1, 64, 0, st>>>(d, 100);
    CK(cudaStreamEndCapture(st, &g));
    cudaGraphExec_t ex; CK(cudaGraphInstantiate(&ex, g, 0));
    auto run_graph = [&] { cudaGraphLaunch(ex, st); cudaStreamSynchronize(st); };
    auto run_eager = [&] {
      reset<<<1, 1, 0, st>>>(ctr); work<<<1, 64, 0, st>>>(d, 100);
      for (int r = 0; r < rounds; ++r) {
        for (int k = 0; k < 5; ++k) work<<<1, 64, 0, st>>>(d, work_n);
        cudaMemcpyAsync(hflag, ctr, 4, cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
      }
      work<<<1, 64, 0, st>>>(d, 100);
      cudaStreamSynchronize(st);
    };
    auto run_eager_nosync = [&] {
      reset<<<1, 1, 0, st>>>(ctr); work<<<1, 64, 0, st>>>(d, 100);
      for (int r = 0; r < rounds; ++r) for (int k = 0; k < 5; ++k) work<<<1, 64, 0, st>>>(d, work_n);
      work<<<1, 64, 0, st>>>(d, 100);
      cudaStreamSynchronize(st);
    };
    auto med = [&](auto f) {
      for (int i = 0; i < 50; ++i) f();
      std::vector<double> t;
      for (int i = 0; i < 500; ++i) {
        auto a = std::chrono::steady_clock::now(); f(); auto b = std::chrono::steady_clock::now();
        t.push_back(std::chrono::duration<double, std::micro>(b - a).count());
      }
      std::sort(t.begin(), t.end()); return t[t.size() / 2];
    };
    const double tg = med(run_graph), te = med(run_eager), tn = med(run_eager_nosync);
    int c = -1; cudaMemcpy(&c, ctr, 4, cudaMemcpyDeviceToHost);
    printf("rounds %d: graph(while) %.1f us, eager sync-per-round %.1f us, eager no-sync %.1f us (counter %d)\n", rounds, tg, te, tn, c);
    cudaGraphExecDestroy(ex); cudaGraphDestroy(g);
  }
  // single-kernel time for scale
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a, st); for (int i = 0; i < 100; ++i) work<<<1, 64, 0, st>>>(d, work_n); cudaEventRecord(b, st); cudaStreamSynchronize(st);
  float ms; cudaEventElapsedTime(&ms, a, b); printf("work kernel %.2f us each (back to back)\n", ms * 10.0);
  return 0;
}
