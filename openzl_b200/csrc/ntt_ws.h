// Device buffers of the NTT path owned by a context (kept apart from ntt.cuh so that the MSM
// translation units, which only need the type inside ozl_ctx, do not recompile when an NTT kernel changes).
#pragma once
#include <stddef.h>

namespace ozl {

// Device tables are kept per direction (index 0 = forward, 1 = inverse) so a prover that
// alternates ifft / coset_fft / coset_ifft on one domain builds each table once.
struct NttWorkspace {
  void* scratch = nullptr; size_t scratch_cap = 0;
  void* tw[2] = {nullptr, nullptr}; size_t tw_cap[2] = {0, 0};
  void* glo[2] = {nullptr, nullptr}; size_t glo_cap[2] = {0, 0};
  void* ghi[2] = {nullptr, nullptr}; size_t ghi_cap[2] = {0, 0};
  void* consts[2] = {nullptr, nullptr};
  int key_field[2] = {-1, -1}, key_log_n[2] = {-1, -1};              // what tw/consts currently hold
  int coset_key_field[2] = {-1, -1}, coset_key_log_n[2] = {-1, -1};
};

}  // namespace ozl
