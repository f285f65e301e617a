// Host-side plumbing shared by every C-ABI entry point: driver entry-point lookup for the tensor-map
// encoder, device queries, version.
#include "common.cuh"
#include <cstdlib>
#include "../../include/pitchextractor_b200.h"

#include <cudaTypedefs.h>
#include <mutex>

namespace pe_host {

static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
static std::once_flag g_encode_once;

static void load_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
      qres == cudaDriverEntryPointSuccess) {
    g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  }
}

int encode_tmap(CUtensorMap* map, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  std::call_once(g_encode_once, load_encode);
  if (!g_encode) return PE_ERR_DRIVER;
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = g_encode(map, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? PE_OK : PE_ERR_DRIVER;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

int check_arch() {
  static int ok = -100;
  if (ok == -100) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return PE_ERR_ARCH;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    ok = (major == 10) ? PE_OK : PE_ERR_ARCH;
  }
  return ok;
}

bool pdl_enabled() {
  static const bool on = []() {
    // measured on the Transformer step (profiles/r02_phase_times.txt): chains of 20-60 us kernels gain ~1 us per
    // boundary, but early-scheduled dependents take SM resources from the second stream's kernels (encoder backward
    // +0.15 ms), so it is opt-in
    const char* e = getenv("PE_PDL");
    return e && e[0] == '1';
  }();
  return on;
}

static salt_addr_fn* salt_table(int** count) {
  static salt_addr_fn fns[8];
  static int n = 0;
  *count = &n;
  return fns;
}
void register_salt(salt_addr_fn fn) {
  int* n;
  salt_addr_fn* fns = salt_table(&n);
  if (*n < 8) fns[(*n)++] = fn;
}

struct SaltAddrs {
  unsigned long long* p[8];
};
__global__ void set_salt_kernel(SaltAddrs a, int n, int slot, unsigned long long salt) {
  if (threadIdx.x < n) a.p[threadIdx.x][slot] = salt;
}

}  // namespace pe_host

extern "C" int pe_set_step_salt(int slot, unsigned long long salt, pe_stream_t stream) {
  if (slot < 0 || slot > 255) return PE_ERR_BAD_SHAPE;
  static pe_host::SaltAddrs addrs;
  static int n_addrs = -1;
  if (n_addrs < 0) {
    int* n;
    pe_host::salt_addr_fn* fns = pe_host::salt_table(&n);
    for (int i = 0; i < *n; ++i) {
      addrs.p[i] = (unsigned long long*)fns[i]();
      if (!addrs.p[i]) return PE_ERR_LAUNCH;
    }
    n_addrs = *n;
  }
  pe_host::set_salt_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(addrs, n_addrs, slot, salt);
  return cudaGetLastError() == cudaSuccess ? PE_OK : PE_ERR_LAUNCH;
}

// Workspace a caller must provide (functions of this library never allocate).  op names the entry point.
extern "C" long long pe_workspace_bytes(const char* op, int B, int T, int L) {
  if (!op || B <= 0) return -1;
  auto is = [&](const char* n) {
    const char* a = op;
    while (*a && *n && *a == *n) { ++a; ++n; }
    return *a == 0 && *n == 0;
  };
  if (is("pe_lstm_seq_fwd") || is("pe_lstm_seq_bwd")) {  // per-step arrival counters of the launch's batch tiles
    if (T <= 0) return -1;
    // (at most num_sms / 24 batch tiles per launch, whatever tile width the launcher picks)
    const int max_bt = pe_host::num_sms() / 24 > 0 ? pe_host::num_sms() / 24 : 6;
    long long bytes = (((long long)4 * max_bt * T * (long long)sizeof(int)) + 255) & ~255LL;
    // backward: + the four transposed recurrent weight matrices of the gate-stacked kernel (narrow batch tiles)
    if (is("pe_lstm_seq_bwd")) bytes += (long long)4 * 1536 * 384 * 2;
    return bytes;
  }
  if (is("pe_logmel_tc")) {  // re-strided waveform copy, only used when rows are not 16-byte aligned
    if (L <= 0) return -1;
    return (long long)B * (((long long)L + 3) / 4 * 4) * (long long)sizeof(float);
  }
  if (is("pe_logmel_f32")) {  // power spectrogram [B*T][n_fft/2+1] with T = frames, L = n_fft here
    if (T <= 0 || L <= 0) return -1;
    return (long long)B * T * (L / 2 + 1) * (long long)sizeof(float);
  }
  return 0;  // every other entry point works in place / in caller-provided outputs
}

extern "C" int pe_version(void) { return 110; }
extern "C" int pe_check_device(void) { return pe_host::check_arch(); }
