// Argument blocks of the row-wise kernels (shared by rowwise.cu and model.cu).
#pragma once
#include "common.cuh"

namespace cgcn {
struct GateFwdArgs {
  const float* y;      // [n][S][D]  A_hat x W + b   (may alias z)
  const float* x;      // [n][S][D]  layer input
  const float* wg;     // [D]        W{1,2}.weight
  const float* bg;     // [1]        W{1,2}.bias
  float* z;            // [n][S][D]  tanh(y)
  float* g;            // [n][S]
  float* xo;           // [n][S][D]  dropout((1-g) x + g z)
  float* stats_partial;// [grid][2][S][D]  sum relu(xo), sum relu(xo)^2   (STATS only)
  int n;
  int gate_off;        // extension: g == 1 (x' = z), the gate parameters are not read
  DropoutCfg drop;
};

struct BnApplyArgs {
  const float* h;
  const float* mean;   // [S][D]
  const float* rstd;   // [S][D]
  const float* gamma;  // [D]
  const float* beta;   // [D]
  float* hb;
  int64_t total4;      // n*S*D/4
  int S, D;
  DropoutCfg drop;
};

struct BnBwdReduceArgs {
  const float* dhb;    // [n][S][D] grad wrt hb
  const float* h;      // [n][S][D] pre-ReLU input of the head
  const float* mean;
  const float* rstd;
  float* partial;      // [grid][2][S][D]: sum dbn, sum dbn*xhat
  int n;
  DropoutCfg drop;
};

struct GateBwdArgs {
  const float* dsrc;   // HEAD: grad wrt hb ; MID: grad wrt the (dropped) layer output
  const float* h;      // HEAD: pre-ReLU head input (= this layer's output)
  const float* mean;   // HEAD [S][D]
  const float* rstd;   // HEAD [S][D]
  const float* gamma;  // HEAD [D]
  const float* c1;     // HEAD [S][D]
  const float* c2;     // HEAD [S][D]
  const float* z;      // [n][S][D]
  const float* x;      // [n][S][D] layer input
  const float* g;      // [n][S]
  const float* wg;     // [D]
  float* dy;           // [n][S][D] grad wrt A_hat x W + b
  float* dxd;          // [n][S][D] (1-g) * dh, or NULL
  float* partial;      // [grid][2*D + 4]: sum dy | sum dgp*z | sum dgp
  int n;
  DropoutCfg drop;     // HEAD: site 1 ; MID: site 0
  const int32_t* scale_rowptr;   // non-NULL: dy is stored as D^-1 dy (1/deg of the window row from this rowptr), the form
                                 // the fused backward layer kernel gathers; the d b column sums stay un-scaled
};

}  // namespace cgcn
