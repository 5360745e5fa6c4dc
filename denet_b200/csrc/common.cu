// Error reporting, device queries and TMA descriptor encoding shared by all C-ABI entry points.
#include <string.h>
#include <atomic>

#include "common.cuh"

namespace dn {

static thread_local char g_err[512] = "";

char* last_error_buf() { return g_err; }

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

static std::atomic<long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
static int g_pdl = 0;     // measured: programmatic edges cost 2 % of the CUDA-graph step (tried and rejected)
int pdl_enabled() { return g_pdl; }
void set_pdl(int on) { g_pdl = on ? 1 : 0; }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

int num_sms() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            cached = 148;
    }
    return cached;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, const uint32_t* estrides) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return set_error(DENET_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0)
        return set_error(DENET_ERR_ARG, "tensor base %p not 16-byte aligned", base);
    cuuint64_t gdim[5];
    cuuint64_t gstr[4];
    cuuint32_t bdim[5];
    cuuint32_t estr[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estr[i] = estrides ? estrides[i] : 1;
    }
    for (int i = 0; i + 1 < rank; ++i) {
        gstr[i] = strides_bytes[i];
        if (gstr[i] % 16 != 0)
            return set_error(DENET_ERR_ARG, "tensor stride %d = %llu bytes is not a multiple of 16", i,
                             (unsigned long long)gstr[i]);
    }
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim,
                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        return set_error(DENET_ERR_CUDA,
                         "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu] box [%u %u %u %u]",
                         (int)r, rank, (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 0),
                         (unsigned long long)(rank > 2 ? gdim[2] : 0), (unsigned long long)(rank > 3 ? gdim[3] : 0),
                         bdim[0], rank > 1 ? bdim[1] : 0, rank > 2 ? bdim[2] : 0, rank > 3 ? bdim[3] : 0);
    }
    return 0;
}

}  // namespace dn

extern "C" const char* denet_last_error(void) { return dn::last_error_buf(); }

extern "C" int denet_abi_version(void) { return DENET_ABI_VERSION; }

namespace dn { long long launch_count(); void set_pdl(int on); }
extern "C" int denet_set_pdl(int on) { dn::set_pdl(on); return 0; }
extern "C" long long denet_launch_count(void) { return dn::launch_count(); }
