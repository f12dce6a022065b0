// probe: programmatic dependent launch (PDL) between kernel nodes inside a WHILE body: does it capture, what does it save
#include <cuda_runtime.h>
#include <cstdio>
#include <chrono>
#include <vector>
#include <algorithm>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("ERR %s line %d: %s\n", #x, __LINE__, cudaGetErrorString(e_)); return 1; } } while (0)
__global__ void work(double* p, int n) {
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  double a = p[threadIdx.x];
  for (int i = 0; i < n; ++i) a = a * 1.0000001 + 1e-9;
  p[threadIdx.x] = a + 1.0;
}
__global__ void cond(int* counter, int rounds, cudaGraphConditionalHandle h) {
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  int c = ++(*counter);
  cudaGraphSetConditional(h, c < rounds ? 1 : 0);
}
__global__ void reset(int* counter) { *counter = 0; }
static bool g_pdl = false;
template <typename... KA, typename... A>
cudaError_t launch(void (*k)(KA...), int grid, int block, cudaStream_t st, A... a) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = g_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, k, KA(a)...);
}
static cudaGraphNode_t leaf_of(cudaGraph_t g) {
  size_t n = 0; cudaGraphGetNodes(g, nullptr, &n); std::vector<cudaGraphNode_t> v(n); cudaGraphGetNodes(g, v.data(), &n);
  cudaGraphNode_t leaf = nullptr;
  for (auto x : v) { size_t nd = 0; cudaGraphNodeGetDependentNodes(x, nullptr, &nd); if (nd == 0) leaf = x; }
  return leaf;
}
int run(bool pdl, int work_n) {
  g_pdl = pdl;
  double* d; int* ctr;
  CK(cudaMalloc(&d, 1024 * 8)); CK(cudaMemset(d, 0, 1024 * 8)); CK(cudaMalloc(&ctr, 4));
  cudaStream_t st; CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  const int rounds = 2;
  cudaGraph_t g; CK(cudaGraphCreate(&g, 0));
  CK(cudaStreamBeginCaptureToGraph(st, g, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
  reset<<<1, 1, 0, st>>>(ctr);
  CK(launch(work, 1, 64, st, d, 100));
  CK(cudaStreamEndCapture(st, &g));
  cudaGraphConditionalHandle h;
  CK(cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault));
  cudaGraphNodeParams p = {}; p.type = cudaGraphNodeTypeConditional; p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
  cudaGraphNode_t leaf = leaf_of(g), cnode;
  CK(cudaGraphAddNode(&cnode, g, &leaf, 1, &p));
  cudaGraph_t body = p.conditional.phGraph_out[0];
  CK(cudaStreamBeginCaptureToGraph(st, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
  for (int k = 0; k < 5; ++k) CK(launch(work, 1, 64, st, d, work_n));
  CK(launch(cond, 1, 1, st, ctr, rounds, h));
  CK(cudaStreamEndCapture(st, &body));
  CK(cudaStreamBeginCaptureToGraph(st, g, &cnode, nullptr, 1, cudaStreamCaptureModeThreadLocal));
  CK(launch(work, 1, 64, st, d, 100));
  CK(cudaStreamEndCapture(st, &g));
  cudaGraphExec_t ex; CK(cudaGraphInstantiate(&ex, g, 0));
  auto f = [&] { cudaGraphLaunch(ex, st); cudaStreamSynchronize(st); };
  for (int i = 0; i < 50; ++i) f();
  std::vector<double> t;
  for (int i = 0; i < 500; ++i) {
    auto a = std::chrono::steady_clock::now(); f(); auto b = std::chrono::steady_clock::now();
    t.push_back(std::chrono::duration<double, std::micro>(b - a).count());
  }
  std::sort(t.begin(), t.end());
  double hd[2]; CK(cudaMemcpy(hd, d, 16, cudaMemcpyDeviceToHost));
  printf("pdl=%d work_n=%d: tick %.1f us (14 kernels), d[0] per tick increments ok=%d\n", (int)pdl, work_n, t[250], hd[0] > 0);
  return 0;
}
int main() {
  for (int n : {1000, 4000}) { if (run(false, n)) return 1; cudaGetLastError(); if (run(true, n)) { printf("PDL variant failed\n"); cudaGetLastError(); } }
  return 0;
}
