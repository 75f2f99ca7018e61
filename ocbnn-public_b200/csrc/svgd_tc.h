// svgd_tc.h — internal interface of the SVGD tensor path (svgd_tc.cu) and the SIMT arm (svgd.cu) used by the C-ABI entry points
// (svgd.cu: ocbnn_svgd_bandwidth / ocbnn_svgd_phi; svgd_step.cu: ocbnn_svgd_step).
#pragma once
#include "common.cuh"

namespace ocbnn {
int svgd_select_mode_value();
int simt_hist_pass(const float* x, long long n, long long d, long long row_start, long long row_step, long long n_rows, int pass, uint32_t* hist,
                   const unsigned long long* state, cudaStream_t st);
int simt_select_update(int pass, uint32_t* hist, unsigned long long* state, long long n, float* out_h, cudaStream_t st);
int simt_phi(const float* x, const float* g, long long n, long long d, long long row_begin, long long n_rows, const float* h, float* out_phi, float* wide,
             cudaStream_t st);
int adagrad_launch(float* x, float* acc, const float* dir, long long n_elem, float lr, float grad_sign, cudaStream_t st);

namespace tc {
long long tc_workspace_bytes(long long n, long long d, long long n_rows);
bool tc_is_flash(long long n, long long d);
int tc_svgd_prepare(const float* x, long long n, long long d, long long row_begin, long long n_rows, void* ws, long long ws_bytes, cudaStream_t st);
int tc_svgd_hist(long long n, long long d, long long row_begin, long long n_rows, int pass, unsigned* hist, const unsigned long long* state, void* ws,
                 cudaStream_t st);
int tc_svgd_select_local(long long n, long long d, void* ws, float* out_h, int mode, unsigned long long* dbg, cudaStream_t st);
int tc_svgd_sel_stage1(long long n, long long d, long long row_begin, long long n_rows, int rank, int world, void* ws, int mode,
                       unsigned long long** red64_out, cudaStream_t st);
int tc_svgd_sel_stage2(long long n, long long d, long long row_begin, long long n_rows, void* ws, cudaStream_t st);
int tc_svgd_sel_hist(long long n, long long d, long long row_begin, long long n_rows, int rank, int world, int pass, void* ws, unsigned** hist_out,
                     cudaStream_t st);
int tc_svgd_sel_update(long long n, long long d, long long row_begin, long long n_rows, int pass, void* ws, float* out_h, cudaStream_t st);
int tc_svgd_phi(const float* g, const float* h, long long n, long long d, long long row_begin, long long n_rows, float* phi, void* ws, cudaStream_t st,
                float* ad_x = nullptr, float* ad_acc = nullptr, float ad_lr = 0.f);
}  // namespace tc
}  // namespace ocbnn
