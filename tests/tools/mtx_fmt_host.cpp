// Host build of the MTX line formatter (stellarscope_b200/csrc/ss_mtx_fmt.cuh) for the CPU unit tests:
// the digit generation is plain integer arithmetic shared by host and device, so it can be checked here
// against Python's exact decimal arithmetic without a GPU.  Test infrastructure only -- the product formats
// on the device (csrc/ss_mtx.cu).
#include "../../stellarscope_b200/csrc/ss_mtx_fmt.cuh"

extern "C" {
int ssmtx_real_supported(double x) { return ssmtx::real_supported(x) ? 1 : 0; }
int ssmtx_put_real(double x, char* out) { return ssmtx::put_real(out, x); }
int ssmtx_put_line(int row, int col, double v, int integer, char* out) {
  return ssmtx::put_line(out, row, col, v, integer != 0);
}
int ssmtx_line_len(int row, int col, double v, int integer) { return ssmtx::line_len(row, col, v, integer != 0); }
int ssmtx_max_line(void) { return ssmtx::MAX_LINE_CHARS; }
}
