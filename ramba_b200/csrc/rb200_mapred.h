// rb200_mapred.h - parameters of the map + reduce kernels (rb200_mapred.cu), filled by the streaming planner
#pragma once
#include "rb200_terms.h"
#include "rb200_vm.cuh"

namespace rb200 {
constexpr int kMrMaxOps = 12;
struct MrParams {
  int mode;            // 0: global reduction, 1: columns
  int src_f32;         // source element type
  const char* src;     // contiguous source
  long long total;     // mode 0: elements
  long long R, C;      // mode 1: rows x columns
  long long rows_per_split;
  int n_chunks, n_split;
  const char* vsrc;    // row-broadcast operand (M_*V), C elements
  int v_f32;
  int n32, n64;        // operations in float32 (source float32 only), then in float64
  int code[kMrMaxOps];
  double w[kMrMaxOps];  // float32 phase: the float value widened (exact)
  int redop;
  KRed red;            // mode 0 output
  u64* red_partials;
  unsigned int* red_counter;
};
struct MrSource {
  const char* base;
  int f32;
  bool row_broadcast;  // column form: stride over the rows is 0
};
int mapred_try(int mode, const TermStep* terms, int n_terms, int n32, const u64* scal, MrSource (*src_of)(void*, const TermStep&), void* ctx, MrParams* out);
cudaError_t mapred_launch(const MrParams& P, unsigned blocks, cudaStream_t stream);
}  // namespace rb200
