// probe: nested WHILE conditional nodes; which graph owns the inner handle; re-arming the inner loop from a kernel
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("ERR %s line %d: %s\n", #x, __LINE__, cudaGetErrorString(e_)); return 1; } } while (0)
__global__ void arm(int* c, cudaGraphConditionalHandle inner) { c[1] = 0; cudaGraphSetConditional(inner, 1); }
__global__ void inner_k(int* c, cudaGraphConditionalHandle inner) { c[2]++; int n = ++c[1]; cudaGraphSetConditional(inner, n < 2 ? 1 : 0); }
__global__ void outer_k(int* c, cudaGraphConditionalHandle outer) { int n = ++c[0]; cudaGraphSetConditional(outer, n < 3 ? 1 : 0); }
__global__ void reset(int* c) { c[0] = c[1] = c[2] = 0; }
static cudaGraphNode_t leaf_of(cudaGraph_t g) {
  size_t n = 0; cudaGraphGetNodes(g, nullptr, &n); std::vector<cudaGraphNode_t> v(n); cudaGraphGetNodes(g, v.data(), &n);
  cudaGraphNode_t leaf = nullptr;
  for (auto x : v) { size_t nd = 0; cudaGraphNodeGetDependentNodes(x, nullptr, &nd); if (nd == 0) leaf = x; }
  return leaf;
}
int run(int owner_is_body) {
  int* c; CK(cudaMalloc(&c, 16));
  cudaStream_t st; CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  cudaGraph_t g; CK(cudaGraphCreate(&g, 0));
  CK(cudaStreamBeginCaptureToGraph(st, g, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
  reset<<<1, 1, 0, st>>>(c);
  CK(cudaStreamEndCapture(st, &g));
  cudaGraphConditionalHandle ho, hi;
  CK(cudaGraphConditionalHandleCreate(&ho, g, 1, cudaGraphCondAssignDefault));
  cudaGraphNodeParams p = {}; p.type = cudaGraphNodeTypeConditional; p.conditional.handle = ho; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
  cudaGraphNode_t leaf = leaf_of(g), on;
  CK(cudaGraphAddNode(&on, g, &leaf, 1, &p));
  cudaGraph_t body = p.conditional.phGraph_out[0];
  CK(cudaGraphConditionalHandleCreate(&hi, owner_is_body ? body : g, 0, 0));
  CK(cudaStreamBeginCaptureToGraph(st, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
  arm<<<1, 1, 0, st>>>(c, hi);
  CK(cudaStreamEndCapture(st, &body));
  cudaGraphNodeParams q = {}; q.type = cudaGraphNodeTypeConditional; q.conditional.handle = hi; q.conditional.type = cudaGraphCondTypeWhile; q.conditional.size = 1;
  cudaGraphNode_t bl = leaf_of(body), in;
  CK(cudaGraphAddNode(&in, body, &bl, 1, &q));
  cudaGraph_t ibody = q.conditional.phGraph_out[0];
  CK(cudaStreamBeginCaptureToGraph(st, ibody, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
  inner_k<<<1, 1, 0, st>>>(c, hi);
  CK(cudaStreamEndCapture(st, &ibody));
  CK(cudaStreamBeginCaptureToGraph(st, body, &in, nullptr, 1, cudaStreamCaptureModeThreadLocal));
  outer_k<<<1, 1, 0, st>>>(c, ho);
  CK(cudaStreamEndCapture(st, &body));
  cudaGraphExec_t ex; CK(cudaGraphInstantiate(&ex, g, 0));
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaGraphLaunch(ex, st)); CK(cudaStreamSynchronize(st));
    int h[3]; CK(cudaMemcpy(h, c, 12, cudaMemcpyDeviceToHost));
    printf("owner_is_body=%d launch %d: outer %d, inner total %d (expect 3, 6)\n", owner_is_body, rep, h[0], h[2]);
  }
  return 0;
}
int main() { int a = run(1); cudaGetLastError(); int b = run(0); return a && b; }
