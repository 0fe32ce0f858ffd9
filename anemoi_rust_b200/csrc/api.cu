// C ABI of libanemoi_b200.so (declared in include/anemoi_b200.h): argument checking that mirrors the
// reference's assert!s, device-buffer management for the host-pointer calls, and the level loop of the
// Jive Merkle builder. All arithmetic happens in the per-field CUDA kernels (field_*.cu); there is no
// CPU implementation of any hash function in this library.
#include <cuda_runtime.h>

#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "anemoi_b200.h"
#include "kernel_args.h"
#include "launch.h"
#include "merkle_aux.h"
#include "nccl_dyn.h"

namespace {

using anemoi::KernelArgs;

thread_local char g_cuda_err[256] = "";

int cuda_fail(cudaError_t e, const char* what) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", what, cudaGetErrorString(e));
    if (e == cudaErrorMemoryAllocation) return ANEMOI_B200_ERR_NOMEM;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return ANEMOI_B200_ERR_NO_DEVICE;
    return ANEMOI_B200_ERR_CUDA;
}
#define CK(call)                                           \
    do {                                                   \
        cudaError_t e__ = (call);                          \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)

typedef cudaError_t (*launch_fn)(int, const KernelArgs*, cudaStream_t);
const launch_fn kLaunch[ANEMOI_NUM_FIELDS] = {
    anemoi_launch_bls12_377, anemoi_launch_bls12_381, anemoi_launch_bn_254, anemoi_launch_ed_on_bls12_377,
    anemoi_launch_jubjub,    anemoi_launch_pallas,    anemoi_launch_vesta,
};
const char* const kFieldName[ANEMOI_NUM_FIELDS] = {"bls12_377", "bls12_381", "bn_254", "ed_on_bls12_377",
                                                   "jubjub",    "pallas",    "vesta"};
const int kFieldLimbs[ANEMOI_NUM_FIELDS] = {6, 6, 4, 4, 4, 4, 4};
// NUM_HASH_ROUNDS: src/<field>/anemoi_2_1/mod.rs:31-32, anemoi_4_3/mod.rs:31-32
const int kRounds[ANEMOI_NUM_FIELDS][2] = {{21, 14}, {21, 14}, {21, 14}, {19, 13}, {21, 14}, {21, 14}, {21, 14}};

// the moduli (curve definitions), u64 limbs little-endian -- only the opt-in input check uses them
const uint64_t kModulus[ANEMOI_NUM_FIELDS][6] = {
    {0x8508c00000000001ULL, 0x170b5d4430000000ULL, 0x1ef3622fba094800ULL, 0x1a22d9f300f5138fULL, 0xc63b05c06ca1493bULL, 0x01ae3a4617c510eaULL},  // bls12_377
    {0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL, 0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL},  // bls12_381
    {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL, 0x0000000000000000ULL, 0x0000000000000000ULL},  // bn_254
    {0x0a11800000000001ULL, 0x59aa76fed0000001ULL, 0x60b44d1e5c37b001ULL, 0x12ab655e9a2ca556ULL, 0x0000000000000000ULL, 0x0000000000000000ULL},  // ed_on_bls12_377
    {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL, 0x0000000000000000ULL, 0x0000000000000000ULL},  // jubjub
    {0x992d30ed00000001ULL, 0x224698fc094cf91bULL, 0x0000000000000000ULL, 0x4000000000000000ULL, 0x0000000000000000ULL, 0x0000000000000000ULL},  // pallas
    {0x8c46eb2100000001ULL, 0x224698fc0994a8ddULL, 0x0000000000000000ULL, 0x4000000000000000ULL, 0x0000000000000000ULL, 0x0000000000000000ULL},  // vesta
};

int check_fi(int field, int inst) {
    if (field < 0 || field >= ANEMOI_NUM_FIELDS) return ANEMOI_B200_ERR_FIELD;
    if (inst != ANEMOI_INST_2_1 && inst != ANEMOI_INST_4_3) return ANEMOI_B200_ERR_INST;
    return ANEMOI_B200_OK;
}

inline int width_of(int inst) { return inst == ANEMOI_INST_2_1 ? 2 : 4; }
inline size_t felt_bytes(int field) { return (size_t)kFieldLimbs[field] * 8; }
inline int aligned16(const void* a, const void* b) {
    return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
}

int launch(int field, int inst, int mode, const void* in, void* out, const uint64_t* offsets, size_t n, size_t len,
           cudaStream_t stream) {
    KernelArgs a;
    a.in = static_cast<const uint32_t*>(in);
    a.out = static_cast<uint32_t*>(out);
    a.offsets = reinterpret_cast<const unsigned long long*>(offsets);
    a.n = n;
    a.len = len;
    a.mode = mode;
    // felts are 32 or 48 bytes, so a 16-byte-aligned base keeps every felt 16-byte aligned
    a.vec16 = aligned16(in, out);
    if (mode == anemoi::MODE_HASH_BYTES || mode == anemoi::MODE_HASH_BYTES_RAGGED)
        a.vec16 = aligned16(out, out);  // the byte input is read bytewise
    const int cols = (mode == anemoi::MODE_TO_BYTES) ? 1 : (inst == ANEMOI_INST_2_1 ? 1 : 2);
    cudaError_t e = kLaunch[field](cols, &a, stream);
    if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
    return ANEMOI_B200_OK;
}

int jive_mode(int inst, int k, int* out_per_state) {
    if (inst == ANEMOI_INST_2_1) {
        if (k != 2) return ANEMOI_B200_ERR_ARITY;  // assert!(k == 2)  anemoi_2_1/hasher.rs:107
        *out_per_state = 1;
        return anemoi::MODE_COMPRESS;
    }
    // assert!(STATE_WIDTH % k == 0); assert!(k % 2 == 0)  anemoi_4_3/hasher.rs:163-165
    if (k == 2) {
        *out_per_state = 2;
        return anemoi::MODE_COMPRESS;
    }
    if (k == 4) {
        *out_per_state = 1;
        return anemoi::MODE_COMPRESS4;
    }
    return ANEMOI_B200_ERR_ARITY;
}

// Device buffers of the host-pointer entry points come from a stream-ordered memory pool OWNED BY THIS LIBRARY
// (one per device, created on first use under std::call_once, release threshold = keep everything), so repeated
// calls reuse the same HBM without cudaMalloc/cudaFree round trips (~10 ms each at 100 MiB) and without touching
// the device's default pool, which other allocators of the process (e.g. PyTorch's) may be using.
// anemoi_b200_pool_trim() hands the cached memory back.
constexpr int kMaxDevices = 64;
std::once_flag g_pool_once[kMaxDevices];
cudaMemPool_t g_pool[kMaxDevices] = {};

cudaMemPool_t library_pool(int device) {
    if (device < 0 || device >= kMaxDevices) return nullptr;
    std::call_once(g_pool_once[device], [device]() {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        cudaMemPool_t pool = nullptr;
        if (cudaMemPoolCreate(&pool, &props) == cudaSuccess) {
            unsigned long long threshold = ~0ULL;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
            g_pool[device] = pool;
        } else {
            cudaGetLastError();  // no pool support: DevBuf falls back to cudaMalloc
        }
    });
    return g_pool[device];
}

struct DevBuf {
    void* p = nullptr;
    cudaStream_t st = nullptr;
    bool async = false;
    ~DevBuf() { release(); }
    void release() {
        if (!p) return;
        if (async) cudaFreeAsync(p, st);
        else cudaFree(p);
        p = nullptr;
    }
    cudaError_t alloc_async(size_t bytes, cudaStream_t stream) {
        int device = 0;
        cudaGetDevice(&device);
        cudaMemPool_t pool = library_pool(device);
        st = stream;
        async = true;
        cudaError_t e = pool ? cudaMallocFromPoolAsync(&p, bytes ? bytes : 16, pool, stream) : cudaErrorNotSupported;
        if (e != cudaSuccess) {  // pool unsupported / exhausted: plain allocation
            cudaGetLastError();
            async = false;
            p = nullptr;
            e = cudaMalloc(&p, bytes ? bytes : 16);
        }
        return e;
    }
};

struct DeviceScope {
    int prev = -1;
    int rc = ANEMOI_B200_OK;
    explicit DeviceScope(int device) {
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0) {
            snprintf(g_cuda_err, sizeof(g_cuda_err), "no CUDA device: %s", cudaGetErrorString(e));
            rc = ANEMOI_B200_ERR_NO_DEVICE;
            return;
        }
        if (device < 0 || device >= count) {
            rc = ANEMOI_B200_ERR_ARG;
            return;
        }
        cudaGetDevice(&prev);
        e = cudaSetDevice(device);
        if (e != cudaSuccess) rc = cuda_fail(e, "cudaSetDevice");
    }
    ~DeviceScope() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// `waiter` does not run past this point before everything enqueued on `producer` so far has finished.
cudaError_t wait_for(cudaStream_t waiter, cudaStream_t producer) {
    cudaEvent_t ev = nullptr;
    cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
    e = cudaEventRecord(ev, producer);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(waiter, ev, 0);
    cudaEventDestroy(ev);  // released by the runtime once the recorded work completes
    return e;
}

// Generic host wrapper: copy `in_bytes` up, run `body(d_in, d_out, stream)`, copy `out_bytes` back.
template <class Body>
int host_call(int device, const void* in, size_t in_bytes, void* out, size_t out_bytes, bool in_place, Body body) {
    DeviceScope scope(device);
    if (scope.rc != ANEMOI_B200_OK) return scope.rc;
    cudaStream_t stream;
    CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    int rc = ANEMOI_B200_OK;
    {
        DevBuf d_in, d_out;  // freed (stream-ordered) before the stream is destroyed
        cudaError_t e = d_in.alloc_async(in_bytes, stream);
        if (e == cudaSuccess && !in_place) e = d_out.alloc_async(out_bytes, stream);
        if (e != cudaSuccess) rc = cuda_fail(e, "device allocation");
        if (rc == ANEMOI_B200_OK) {
            e = cudaMemcpyAsync(d_in.p, in, in_bytes, cudaMemcpyHostToDevice, stream);
            if (e != cudaSuccess) rc = cuda_fail(e, "H2D copy");
        }
        if (rc == ANEMOI_B200_OK) rc = body(d_in.p, in_place ? d_in.p : d_out.p, stream);
        if (rc == ANEMOI_B200_OK) {
            e = cudaMemcpyAsync(out, in_place ? d_in.p : d_out.p, out_bytes, cudaMemcpyDeviceToHost, stream);
            if (e != cudaSuccess) rc = cuda_fail(e, "D2H copy");
        }
        e = cudaStreamSynchronize(stream);
        if (rc == ANEMOI_B200_OK && e != cudaSuccess) rc = cuda_fail(e, "cudaStreamSynchronize");
    }
    cudaStreamSynchronize(stream);
    cudaStreamDestroy(stream);
    return rc;
}

// Host wrapper for big batches of INDEPENDENT units (Jive compress): the batch goes up, through the kernel and back in
// `chunks` pieces on separate streams -- upload of piece c+1 and download of piece c-1 overlap the hashing of piece c,
// and the pieces alternate between two compute streams so that the next one starts filling SMs while the previous one
// drains (one stream would pay a partial last wave per piece). body(d_in, d_out, n_units, stream) enqueues one piece.
template <class Body>
int host_call_chunked(int device, const void* in, size_t in_unit, void* out, size_t out_unit, size_t n, int chunks, Body body) {
    DeviceScope scope(device);
    if (scope.rc != ANEMOI_B200_OK) return scope.rc;
    cudaStream_t s_main = nullptr, s_in = nullptr, s_out = nullptr, s_cmp[2] = {nullptr, nullptr};
    int rc = ANEMOI_B200_OK;
    auto fail = [&](cudaError_t e, const char* what) {
        if (e != cudaSuccess && rc == ANEMOI_B200_OK) rc = cuda_fail(e, what);
        return e != cudaSuccess;
    };
    fail(cudaStreamCreateWithFlags(&s_main, cudaStreamNonBlocking), "cudaStreamCreate");
    if (!rc) fail(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking), "cudaStreamCreate");
    if (!rc) fail(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking), "cudaStreamCreate");
    for (int i = 0; i < 2 && !rc; i++) fail(cudaStreamCreateWithFlags(&s_cmp[i], cudaStreamNonBlocking), "cudaStreamCreate");
    {
        DevBuf d_in, d_out;  // allocated and freed on s_main, which is idle in between
        if (!rc) fail(d_in.alloc_async(n * in_unit, s_main), "device allocation");
        if (!rc) fail(d_out.alloc_async(n * out_unit, s_main), "device allocation");
        if (!rc) fail(cudaStreamSynchronize(s_main), "device allocation");  // the buffers exist before the other streams use them
        const uint8_t* h_in = static_cast<const uint8_t*>(in);
        uint8_t* h_out = static_cast<uint8_t*>(out);
        for (int c = 0; c < chunks && !rc; c++) {
            const size_t lo = n * (size_t)c / (size_t)chunks, hi = n * (size_t)(c + 1) / (size_t)chunks;
            if (hi == lo) continue;
            uint8_t* di = static_cast<uint8_t*>(d_in.p) + lo * in_unit;
            uint8_t* dout = static_cast<uint8_t*>(d_out.p) + lo * out_unit;
            cudaStream_t sc = s_cmp[c & 1];
            if (fail(cudaMemcpyAsync(di, h_in + lo * in_unit, (hi - lo) * in_unit, cudaMemcpyHostToDevice, s_in), "H2D copy")) break;
            if (fail(wait_for(sc, s_in), "stream ordering")) break;
            rc = body(di, dout, hi - lo, sc);
            if (rc) break;
            if (fail(wait_for(s_out, sc), "stream ordering")) break;
            fail(cudaMemcpyAsync(h_out + lo * out_unit, dout, (hi - lo) * out_unit, cudaMemcpyDeviceToHost, s_out), "D2H copy");
        }
        // everything must have finished before the buffers are released (and before the caller reads `out`)
        cudaStream_t all[4] = {s_in, s_cmp[0], s_cmp[1], s_out};
        for (cudaStream_t st : all)
            if (st) fail(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    }
    cudaStream_t all[5] = {s_main, s_in, s_cmp[0], s_cmp[1], s_out};
    for (cudaStream_t st : all)
        if (st) {
            cudaStreamSynchronize(st);
            cudaStreamDestroy(st);
        }
    return rc;
}

int merkle_check(int field, int inst, int arity) {
    int rc = check_fi(field, inst);
    if (rc) return rc;
    if (!((inst == ANEMOI_INST_2_1 && arity == 2) || (inst == ANEMOI_INST_4_3 && arity == 4))) return ANEMOI_B200_ERR_ARITY;
    return ANEMOI_B200_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
extern "C" {

int anemoi_b200_version(void) { return ANEMOI_B200_VERSION; }

const char* anemoi_b200_strerror(int code) {
    switch (code) {
        case ANEMOI_B200_OK: return "ok";
        case ANEMOI_B200_ERR_ARG: return "bad argument (null pointer, device index or variant)";
        case ANEMOI_B200_ERR_FIELD: return "unknown field id";
        case ANEMOI_B200_ERR_INST: return "unknown instantiation id";
        case ANEMOI_B200_ERR_ARITY: return "compression factor / arity not supported by this instantiation";
        case ANEMOI_B200_ERR_LENGTH: return "length is not a whole number of states or not a power of the arity";
        case ANEMOI_B200_ERR_CUDA: return "CUDA runtime error (see anemoi_b200_last_cuda_error)";
        case ANEMOI_B200_ERR_NO_DEVICE: return "no CUDA device available (this library has no CPU fallback)";
        case ANEMOI_B200_ERR_NOMEM: return "device memory allocation failed";
        case ANEMOI_B200_ERR_NCCL: return "NCCL error or libnccl.so.2 not loadable (see anemoi_b200_last_cuda_error)";
    }
    return "unknown error code";
}

const char* anemoi_b200_last_cuda_error(void) { return g_cuda_err; }

int anemoi_b200_pool_trim(int device, size_t keep_bytes) {
    DeviceScope scope(device);
    if (scope.rc != ANEMOI_B200_OK) return scope.rc;
    cudaMemPool_t pool = library_pool(device);
    if (!pool) return ANEMOI_B200_OK;
    CK(cudaDeviceSynchronize());
    CK(cudaMemPoolTrimTo(pool, keep_bytes));
    return ANEMOI_B200_OK;
}

int anemoi_b200_pool_reserve(int device, size_t bytes) {
    DeviceScope scope(device);
    if (scope.rc != ANEMOI_B200_OK) return scope.rc;
    if (bytes == 0) return ANEMOI_B200_OK;
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    int rc = ANEMOI_B200_OK;
    {
        DevBuf block;  // one allocation of the whole size, handed straight back: the pool keeps the memory
        cudaError_t e = block.alloc_async(bytes, st);
        if (e != cudaSuccess) rc = cuda_fail(e, "device allocation");
    }
    cudaError_t e = cudaStreamSynchronize(st);
    if (!rc && e != cudaSuccess) rc = cuda_fail(e, "cudaStreamSynchronize");
    cudaStreamDestroy(st);
    return rc;
}

int anemoi_b200_device_count(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return count;
}

int anemoi_b200_field_limbs(int field) {
    if (field < 0 || field >= ANEMOI_NUM_FIELDS) return ANEMOI_B200_ERR_FIELD;
    return kFieldLimbs[field];
}
int anemoi_b200_state_width(int inst) {
    if (inst != ANEMOI_INST_2_1 && inst != ANEMOI_INST_4_3) return ANEMOI_B200_ERR_INST;
    return width_of(inst);
}
int anemoi_b200_rate_width(int inst) {
    if (inst != ANEMOI_INST_2_1 && inst != ANEMOI_INST_4_3) return ANEMOI_B200_ERR_INST;
    return inst == ANEMOI_INST_2_1 ? 1 : 3;
}
int anemoi_b200_num_rounds(int field, int inst) {
    int rc = check_fi(field, inst);
    if (rc) return rc;
    return kRounds[field][inst];
}
const char* anemoi_b200_field_name(int field) {
    if (field < 0 || field >= ANEMOI_NUM_FIELDS) return nullptr;
    return kFieldName[field];
}

// ---- device-pointer entry points ---------------------------------------------------------------

int anemoi_b200_count_noncanonical_dev(int field, const uint64_t* d_elems, size_t n, uint64_t* d_count, void* stream) {
    int rc = check_fi(field, ANEMOI_INST_2_1);
    if (rc) return rc;
    if (!d_count || (n && !d_elems)) return ANEMOI_B200_ERR_ARG;
    CK(anemoi_aux_count_noncanonical(d_elems, n, kFieldLimbs[field], kModulus[field], (unsigned long long*)d_count,
                                     (cudaStream_t)stream));
    return ANEMOI_B200_OK;
}

int anemoi_b200_permute_dev(int field, int inst, uint64_t* d_states, size_t n, void* stream) {
    int rc = check_fi(field, inst);
    if (rc) return rc;
    if (n == 0) return ANEMOI_B200_OK;
    if (!d_states) return ANEMOI_B200_ERR_ARG;
    return launch(field, inst, anemoi::MODE_PERMUTE, d_states, d_states, nullptr, n, 0, (cudaStream_t)stream);
}

int anemoi_b200_sbox_layer_dev(int field, int inst, uint64_t* d_states, size_t n, void* stream) {
    int rc = check_fi(field, inst);
    if (rc) return rc;
    if (n == 0) return ANEMOI_B200_OK;
    if (!d_states) return ANEMOI_B200_ERR_ARG;
    return launch(field, inst, anemoi::MODE_SBOX, d_states, d_states, nullptr, n, 0, (cudaStream_t)stream);
}

int anemoi_b200_layer_dev(int field, int inst, int layer, int round, uint64_t* d_states, size_t n, void* stream) {
    int rc = check_fi(field, inst);
    if (rc) return rc;
    if (layer < 0 || layer > 3) return ANEMOI_B200_ERR_ARG;
    // assert!(round_ctr < Self::NUM_ROUNDS)  src/traits.rs:115
    if ((layer == 0 || layer == 3) && (round < 0 || round >= kRounds[field][inst])) return ANEMOI_B200_ERR_LENGTH;
    if (n == 0) return ANEMOI_B200_OK;
    if (!d_states) return ANEMOI_B200_ERR_ARG;
    const int modes[4] = {anemoi::MODE_LAYER_ARK, anemoi::MODE_LAYER_MDS, anemoi::MODE_SBOX, anemoi::MODE_LAYER_ROUND};
    return launch(field, inst, modes[layer], d_states, d_states, nullptr, n, (size_t)(round < 0 ? 0 : round), (cudaStream_t)stream);
}

int anemoi_b200_compress_dev(int field, int inst, int k, const uint64_t* d_in, uint64_t* d_out, size_t n, void* stream) {
    int rc = check_fi(field, inst);
    if (rc) return rc;
    int per = 0;
    int mode = jive_mode(inst, k, &per);
    if (mode < 0) return mode;
    if (n == 0) return ANEMOI_B200_OK;
    if (!d_in || !d_out) return ANEMOI_B200_ERR_ARG;
    return launch(field, inst, mode, d_in, d_out, nullptr, n, 0, (cudaStream_t)stream);
}

int anemoi_b200_hash_field_dev(int field, int inst, const uint64_t* d_elems, size_t n_msgs, size_t felts_per_msg,
                               uint64_t* d_digests, void* stream) {
    int rc = check_fi(field, inst);
    if (rc) return rc;
    if (n_msgs == 0) return ANEMOI_B200_OK;
    if (!d_digests || (!d_elems && felts_per_msg)) return ANEMOI_B200_ERR_ARG;
    return launch(field, inst, anemoi::MODE_HASH, d_elems, d_digests, nullptr, n_msgs, felts_per_msg, (cudaStream_t)stream);
}

int anemoi_b200_hash_field_ragged_dev(int field, int inst, const uint64_t* d_elems, const uint64_t* d_offsets,
                                      size_t n_msgs, uint64_t* d_digests, void* stream) {
    int rc = check_fi(field, inst);
    if (rc) return rc;
    if (n_msgs == 0) return ANEMOI_B200_OK;
    if (!d_digests || !d_offsets) return ANEMOI_B200_ERR_ARG;
    return launch(field, inst, anemoi::MODE_HASH_RAGGED, d_elems, d_digests, d_offsets, n_msgs, 0, (cudaStream_t)stream);
}

int anemoi_b200_hash_bytes_dev(int field, int inst, const uint8_t* d_bytes, size_t n_msgs, size_t bytes_per_msg,
                               uint64_t* d_digests, void* stream) {
    int rc = check_fi(field, inst);
    if (rc) return rc;
    if (n_msgs == 0) return ANEMOI_B200_OK;
    if (!d_digests || (!d_bytes && bytes_per_msg)) return ANEMOI_B200_ERR_ARG;
    return launch(field, inst, anemoi::MODE_HASH_BYTES, d_bytes, d_digests, nullptr, n_msgs, bytes_per_msg,
                  (cudaStream_t)stream);
}

int anemoi_b200_hash_bytes_ragged_dev(int field, int inst, const uint8_t* d_bytes, const uint64_t* d_offsets,
                                      size_t n_msgs, uint64_t* d_digests, void* stream) {
    int rc = check_fi(field, inst);
    if (rc) return rc;
    if (n_msgs == 0) return ANEMOI_B200_OK;
    if (!d_digests || !d_offsets) return ANEMOI_B200_ERR_ARG;
    return launch(field, inst, anemoi::MODE_HASH_BYTES_RAGGED, d_bytes, d_digests, d_offsets, n_msgs, 0, (cudaStream_t)stream);
}

int anemoi_b200_merge_dev(int field, int inst, const uint64_t* d_pairs, uint64_t* d_out, size_t n, void* stream) {
    int rc = check_fi(field, inst);
    if (rc) return rc;
    if (n == 0) return ANEMOI_B200_OK;
    if (!d_pairs || !d_out) return ANEMOI_B200_ERR_ARG;
    const int mode = inst == ANEMOI_INST_2_1 ? anemoi::MODE_COMPRESS : anemoi::MODE_MERGE43;
    return launch(field, inst, mode, d_pairs, d_out, nullptr, n, 0, (cudaStream_t)stream);
}

int anemoi_b200_digest_to_bytes_dev(int field, const uint64_t* d_digests, uint8_t* d_bytes, size_t n, void* stream) {
    int rc = check_fi(field, ANEMOI_INST_2_1);
    if (rc) return rc;
    if (n == 0) return ANEMOI_B200_OK;
    if (!d_digests || !d_bytes) return ANEMOI_B200_ERR_ARG;
    if (reinterpret_cast<uintptr_t>(d_bytes) & 7) return ANEMOI_B200_ERR_ARG;
    return launch(field, ANEMOI_INST_2_1, anemoi::MODE_TO_BYTES, d_digests, d_bytes, nullptr, n, 0, (cudaStream_t)stream);
}

size_t anemoi_b200_merkle_scratch_felts(int arity, size_t n_leaves) {
    if (arity < 2) return 0;
    const size_t l1 = n_leaves / (size_t)arity;
    return l1 + l1 / (size_t)arity + 2;
}

int anemoi_b200_merkle_reduce_dev(int field, int inst, int arity, const uint64_t* d_leaves, size_t n_leaves, int levels,
                                  uint64_t* d_scratch, uint64_t* d_out, void* stream) {
    int rc = merkle_check(field, inst, arity);
    if (rc) return rc;
    if (levels < 0) return ANEMOI_B200_ERR_ARG;
    if (n_leaves == 0) return ANEMOI_B200_ERR_LENGTH;
    size_t div = 1;
    for (int l = 0; l < levels; l++) {
        if (div > n_leaves / (size_t)arity) return ANEMOI_B200_ERR_LENGTH;
        div *= (size_t)arity;
    }
    if (n_leaves % div != 0) return ANEMOI_B200_ERR_LENGTH;
    if (!d_leaves || !d_out) return ANEMOI_B200_ERR_ARG;
    if (levels > 1 && !d_scratch) return ANEMOI_B200_ERR_ARG;
    const size_t fb = felt_bytes(field);
    cudaStream_t st = (cudaStream_t)stream;
    if (levels == 0) {
        CK(cudaMemcpyAsync(d_out, d_leaves, n_leaves * fb, cudaMemcpyDeviceToDevice, st));
        return ANEMOI_B200_OK;
    }
    const int mode = arity == 2 ? anemoi::MODE_COMPRESS : anemoi::MODE_COMPRESS4;
    const size_t words = (size_t)kFieldLimbs[field];
    uint64_t* ping = d_scratch;                                      // n/arity felts
    uint64_t* pong = d_scratch ? d_scratch + (n_leaves / arity + 1) * words : nullptr;  // n/arity^2 felts
    const uint64_t* src = d_leaves;
    size_t n = n_leaves;
    for (int l = 0; l < levels; l++) {
        const size_t nodes = n / (size_t)arity;
        uint64_t* dst = (l == levels - 1) ? d_out : ((l & 1) ? pong : ping);
        rc = launch(field, inst, mode, src, dst, nullptr, nodes, 0, st);
        if (rc) return rc;
        src = dst;
        n = nodes;
    }
    return ANEMOI_B200_OK;
}

// ---- retained trees, openings, verification ------------------------------------------------------

size_t anemoi_b200_merkle_tree_felts(int arity, size_t n_leaves) {
    if (arity < 2 || n_leaves == 0) return 0;
    return (n_leaves - 1) / (size_t)(arity - 1);
}

static int tree_height(int arity, size_t n_leaves, int* height) {
    if (n_leaves == 0) return ANEMOI_B200_ERR_LENGTH;
    int h = 0;
    size_t m = n_leaves;
    while (m > 1) {
        if (m % (size_t)arity) return ANEMOI_B200_ERR_LENGTH;
        m /= (size_t)arity;
        h++;
    }
    *height = h;
    return ANEMOI_B200_OK;
}

int anemoi_b200_merkle_tree_dev(int field, int inst, int arity, const uint64_t* d_leaves, size_t n_leaves,
                                uint64_t* d_tree, void* stream) {
    int rc = merkle_check(field, inst, arity);
    if (rc) return rc;
    int height = 0;
    rc = tree_height(arity, n_leaves, &height);
    if (rc) return rc;
    if (height == 0) return ANEMOI_B200_OK;  // a single leaf is its own root; the tree above it is empty
    if (!d_leaves || !d_tree) return ANEMOI_B200_ERR_ARG;
    const int mode = arity == 2 ? anemoi::MODE_COMPRESS : anemoi::MODE_COMPRESS4;
    const size_t words = (size_t)kFieldLimbs[field];
    const uint64_t* src = d_leaves;
    uint64_t* dst = d_tree;
    size_t n = n_leaves;
    for (int l = 0; l < height; l++) {
        const size_t nodes = n / (size_t)arity;
        rc = launch(field, inst, mode, src, dst, nullptr, nodes, 0, (cudaStream_t)stream);
        if (rc) return rc;
        src = dst;
        dst += nodes * words;
        n = nodes;
    }
    return ANEMOI_B200_OK;
}

int anemoi_b200_merkle_open_dev(int field, int inst, int arity, const uint64_t* d_leaves, const uint64_t* d_tree,
                                size_t n_leaves, const uint64_t* d_indices, size_t n_idx, uint64_t* d_paths, void* stream) {
    int rc = merkle_check(field, inst, arity);
    if (rc) return rc;
    int height = 0;
    rc = tree_height(arity, n_leaves, &height);
    if (rc) return rc;
    if (n_idx == 0 || height == 0) return ANEMOI_B200_OK;
    if (!d_leaves || !d_tree || !d_indices || !d_paths) return ANEMOI_B200_ERR_ARG;
    CK(anemoi_aux_gather_paths(d_leaves, d_tree, n_leaves, arity, height, kFieldLimbs[field], d_indices, n_idx, d_paths,
                               (cudaStream_t)stream));
    return ANEMOI_B200_OK;
}

int anemoi_b200_merkle_verify_dev(int field, int inst, int arity, const uint64_t* d_leaf_values, const uint64_t* d_indices,
                                  const uint64_t* d_paths, int height, size_t n_idx, uint64_t* d_scratch, uint64_t* d_roots,
                                  void* stream) {
    int rc = merkle_check(field, inst, arity);
    if (rc) return rc;
    if (height < 0) return ANEMOI_B200_ERR_ARG;
    if (n_idx == 0) return ANEMOI_B200_OK;
    if (!d_leaf_values || !d_roots) return ANEMOI_B200_ERR_ARG;
    const size_t fb = felt_bytes(field);
    cudaStream_t st = (cudaStream_t)stream;
    if (height == 0) {
        CK(cudaMemcpyAsync(d_roots, d_leaf_values, n_idx * fb, cudaMemcpyDeviceToDevice, st));
        return ANEMOI_B200_OK;
    }
    if (!d_indices || !d_paths || !d_scratch) return ANEMOI_B200_ERR_ARG;
    const int mode = arity == 2 ? anemoi::MODE_COMPRESS : anemoi::MODE_COMPRESS4;
    const size_t words = (size_t)kFieldLimbs[field];
    uint64_t* states = d_scratch;                          // n_idx * arity felts
    uint64_t* cur = d_scratch + n_idx * (size_t)arity * words;  // n_idx felts
    const uint64_t* src = d_leaf_values;
    for (int l = 0; l < height; l++) {
        CK(anemoi_aux_assemble_level(src, d_paths, d_indices, l, arity, height, (int)words, n_idx, states, st));
        uint64_t* dst = (l == height - 1) ? d_roots : cur;
        rc = launch(field, inst, mode, states, dst, nullptr, n_idx, 0, st);
        if (rc) return rc;
        src = dst;
    }
    return ANEMOI_B200_OK;
}

// ---- host-pointer entry points -----------------------------------------------------------------

int anemoi_b200_permute(int field, int inst, uint64_t* states, size_t n, int device) {
    int rc = check_fi(field, inst);
    if (rc) return rc;
    if (n == 0) return ANEMOI_B200_OK;
    if (!states) return ANEMOI_B200_ERR_ARG;
    const size_t bytes = n * width_of(inst) * felt_bytes(field);
    return host_call(device, states, bytes, states, bytes, true, [&](void* di, void*, cudaStream_t st) {
        return anemoi_b200_permute_dev(field, inst, (uint64_t*)di, n, st);
    });
}

int anemoi_b200_sbox_layer(int field, int inst, uint64_t* states, size_t n, int device) {
    int rc = check_fi(field, inst);
    if (rc) return rc;
    if (n == 0) return ANEMOI_B200_OK;
    if (!states) return ANEMOI_B200_ERR_ARG;
    const size_t bytes = n * width_of(inst) * felt_bytes(field);
    return host_call(device, states, bytes, states, bytes, true, [&](void* di, void*, cudaStream_t st) {
        return anemoi_b200_sbox_layer_dev(field, inst, (uint64_t*)di, n, st);
    });
}

int anemoi_b200_layer(int field, int inst, int layer, int round, uint64_t* states, size_t n, int device) {
    int rc = check_fi(field, inst);
    if (rc) return rc;
    if (layer < 0 || layer > 3) return ANEMOI_B200_ERR_ARG;
    if ((layer == 0 || layer == 3) && (round < 0 || round >= kRounds[field][inst])) return ANEMOI_B200_ERR_LENGTH;
    if (n == 0) return ANEMOI_B200_OK;
    if (!states) return ANEMOI_B200_ERR_ARG;
    const size_t bytes = n * width_of(inst) * felt_bytes(field);
    return host_call(device, states, bytes, states, bytes, true, [&](void* di, void*, cudaStream_t st) {
        return anemoi_b200_layer_dev(field, inst, layer, round, (uint64_t*)di, n, st);
    });
}

int anemoi_b200_compress(int field, int inst, int k, const uint64_t* in, uint64_t* out, size_t n, int device) {
    int rc = check_fi(field, inst);
    if (rc) return rc;
    int per = 0;
    int mode = jive_mode(inst, k, &per);
    if (mode < 0) return mode;
    if (n == 0) return ANEMOI_B200_OK;
    if (!in || !out) return ANEMOI_B200_ERR_ARG;
    const size_t fb = felt_bytes(field);
    if (n >= ((size_t)1 << 19))  // big batch: pipeline upload / hashing / download in 4 pieces
        return host_call_chunked(device, in, width_of(inst) * fb, out, per * fb, n, 4,
                                 [&](void* di, void* dout, size_t m, cudaStream_t st) {
                                     return anemoi_b200_compress_dev(field, inst, k, (const uint64_t*)di, (uint64_t*)dout, m, st);
                                 });
    return host_call(device, in, n * width_of(inst) * fb, out, n * per * fb, false, [&](void* di, void* dout, cudaStream_t st) {
        return anemoi_b200_compress_dev(field, inst, k, (const uint64_t*)di, (uint64_t*)dout, n, st);
    });
}

int anemoi_b200_compress_multi(int field, int inst, int k, const uint64_t* in, uint64_t* out, size_t n, int n_gpus) {
    int rc = check_fi(field, inst);
    if (rc) return rc;
    int per = 0;
    int mode = jive_mode(inst, k, &per);
    if (mode < 0) return mode;
    if (n == 0) return ANEMOI_B200_OK;
    if (!in || !out) return ANEMOI_B200_ERR_ARG;
    const int count = anemoi_b200_device_count();
    if (count == 0) {
        snprintf(g_cuda_err, sizeof(g_cuda_err), "no CUDA device");
        return ANEMOI_B200_ERR_NO_DEVICE;
    }
    if (n_gpus < 1 || n_gpus > count) return ANEMOI_B200_ERR_ARG;
    if ((size_t)n_gpus > n) n_gpus = (int)n;
    if (n_gpus == 1) return anemoi_b200_compress(field, inst, k, in, out, n, 0);
    const size_t in_words = (size_t)width_of(inst) * kFieldLimbs[field], out_words = (size_t)per * kFieldLimbs[field];
    std::vector<int> rcs(n_gpus, ANEMOI_B200_OK);
    std::vector<std::string> errs(n_gpus);
    std::vector<std::thread> th;
    for (int g = 0; g < n_gpus; g++) {
        const size_t lo = n * (size_t)g / (size_t)n_gpus, hi = n * (size_t)(g + 1) / (size_t)n_gpus;
        th.emplace_back([&, g, lo, hi]() {
            rcs[g] = anemoi_b200_compress(field, inst, k, in + lo * in_words, out + lo * out_words, hi - lo, g);
            errs[g] = g_cuda_err;
        });
    }
    for (auto& t : th) t.join();
    for (int g = 0; g < n_gpus; g++)
        if (rcs[g]) {
            snprintf(g_cuda_err, sizeof(g_cuda_err), "gpu %d: %s", g, errs[g].c_str());
            return rcs[g];
        }
    return ANEMOI_B200_OK;
}

int anemoi_b200_hash_field(int field, int inst, const uint64_t* elems, size_t n_msgs, size_t felts_per_msg,
                           uint64_t* digests, int device) {
    int rc = check_fi(field, inst);
    if (rc) return rc;
    if (n_msgs == 0) return ANEMOI_B200_OK;
    if (!digests || (!elems && felts_per_msg)) return ANEMOI_B200_ERR_ARG;
    const size_t fb = felt_bytes(field);
    return host_call(device, elems, n_msgs * felts_per_msg * fb, digests, n_msgs * fb, false,
                     [&](void* di, void* dout, cudaStream_t st) {
                         return anemoi_b200_hash_field_dev(field, inst, (const uint64_t*)di, n_msgs, felts_per_msg,
                                                           (uint64_t*)dout, st);
                     });
}

int anemoi_b200_hash_field_ragged(int field, int inst, const uint64_t* elems, const uint64_t* offsets, size_t n_msgs,
                                  uint64_t* digests, int device) {
    int rc = check_fi(field, inst);
    if (rc) return rc;
    if (n_msgs == 0) return ANEMOI_B200_OK;
    if (!digests || !offsets) return ANEMOI_B200_ERR_ARG;
    for (size_t i = 0; i < n_msgs; i++)
        if (offsets[i + 1] < offsets[i]) return ANEMOI_B200_ERR_LENGTH;
    const uint64_t base = offsets[0], total = offsets[n_msgs] - base;
    if (total && !elems) return ANEMOI_B200_ERR_ARG;
    const size_t fb = felt_bytes(field);
    // rebase the offsets so only the referenced element range is copied
    std::vector<uint64_t> rel(n_msgs + 1);
    for (size_t i = 0; i <= n_msgs; i++) rel[i] = offsets[i] - base;
    const uint8_t* src = reinterpret_cast<const uint8_t*>(elems) + base * fb;
    return host_call(device, total ? src : reinterpret_cast<const uint8_t*>(rel.data()), total * fb, digests, n_msgs * fb,
                     false, [&](void* di, void* dout, cudaStream_t st) {
                         // the offsets go up on the SAME stream as the kernel that reads them (stream-ordered)
                         DevBuf d_off;
                         CK(d_off.alloc_async((n_msgs + 1) * sizeof(uint64_t), st));
                         CK(cudaMemcpyAsync(d_off.p, rel.data(), (n_msgs + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
                         return anemoi_b200_hash_field_ragged_dev(field, inst, (const uint64_t*)di,
                                                                  (const uint64_t*)d_off.p, n_msgs, (uint64_t*)dout, st);
                     });
}

int anemoi_b200_hash_bytes(int field, int inst, const uint8_t* bytes, size_t n_msgs, size_t bytes_per_msg,
                           uint64_t* digests, int device) {
    int rc = check_fi(field, inst);
    if (rc) return rc;
    if (n_msgs == 0) return ANEMOI_B200_OK;
    if (!digests || (!bytes && bytes_per_msg)) return ANEMOI_B200_ERR_ARG;
    const size_t fb = felt_bytes(field);
    return host_call(device, bytes_per_msg ? (const void*)bytes : (const void*)digests, n_msgs * bytes_per_msg, digests,
                     n_msgs * fb, false, [&](void* di, void* dout, cudaStream_t st) {
                         return anemoi_b200_hash_bytes_dev(field, inst, (const uint8_t*)di, n_msgs, bytes_per_msg,
                                                           (uint64_t*)dout, st);
                     });
}

int anemoi_b200_hash_bytes_ragged(int field, int inst, const uint8_t* bytes, const uint64_t* offsets, size_t n_msgs,
                                  uint64_t* digests, int device) {
    int rc = check_fi(field, inst);
    if (rc) return rc;
    if (n_msgs == 0) return ANEMOI_B200_OK;
    if (!digests || !offsets) return ANEMOI_B200_ERR_ARG;
    for (size_t i = 0; i < n_msgs; i++)
        if (offsets[i + 1] < offsets[i]) return ANEMOI_B200_ERR_LENGTH;
    const uint64_t base = offsets[0], total = offsets[n_msgs] - base;
    if (total && !bytes) return ANEMOI_B200_ERR_ARG;
    std::vector<uint64_t> rel(n_msgs + 1);
    for (size_t i = 0; i <= n_msgs; i++) rel[i] = offsets[i] - base;
    const void* src = total ? (const void*)(bytes + base) : (const void*)rel.data();
    return host_call(device, src, total, digests, n_msgs * felt_bytes(field), false, [&](void* di, void* dout, cudaStream_t st) {
        DevBuf d_off;  // same stream as the kernel that reads it
        CK(d_off.alloc_async((n_msgs + 1) * sizeof(uint64_t), st));
        CK(cudaMemcpyAsync(d_off.p, rel.data(), (n_msgs + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        return anemoi_b200_hash_bytes_ragged_dev(field, inst, (const uint8_t*)di, (const uint64_t*)d_off.p, n_msgs,
                                                 (uint64_t*)dout, st);
    });
}

int anemoi_b200_merge(int field, int inst, const uint64_t* digest_pairs, uint64_t* out, size_t n, int device) {
    int rc = check_fi(field, inst);
    if (rc) return rc;
    if (n == 0) return ANEMOI_B200_OK;
    if (!digest_pairs || !out) return ANEMOI_B200_ERR_ARG;
    const size_t fb = felt_bytes(field);
    return host_call(device, digest_pairs, n * 2 * fb, out, n * fb, false, [&](void* di, void* dout, cudaStream_t st) {
        return anemoi_b200_merge_dev(field, inst, (const uint64_t*)di, (uint64_t*)dout, n, st);
    });
}

int anemoi_b200_count_noncanonical(int field, const uint64_t* elems, size_t n, uint64_t* count, int device) {
    int rc = check_fi(field, ANEMOI_INST_2_1);
    if (rc) return rc;
    if (!count || (n && !elems)) return ANEMOI_B200_ERR_ARG;
    if (n == 0) {
        *count = 0;
        return ANEMOI_B200_OK;
    }
    return host_call(device, elems, n * felt_bytes(field), count, sizeof(uint64_t), false, [&](void* di, void* dout, cudaStream_t st) {
        return anemoi_b200_count_noncanonical_dev(field, (const uint64_t*)di, n, (uint64_t*)dout, st);
    });
}

int anemoi_b200_digest_to_bytes(int field, const uint64_t* digests, uint8_t* bytes, size_t n, int device) {
    int rc = check_fi(field, ANEMOI_INST_2_1);
    if (rc) return rc;
    if (n == 0) return ANEMOI_B200_OK;
    if (!digests || !bytes) return ANEMOI_B200_ERR_ARG;
    const size_t fb = felt_bytes(field);
    return host_call(device, digests, n * fb, bytes, n * fb, false, [&](void* di, void* dout, cudaStream_t st) {
        return anemoi_b200_digest_to_bytes_dev(field, (const uint64_t*)di, (uint8_t*)dout, n, st);
    });
}

// ---- sharded Merkle root: per-rank sub-tree, ONE NCCL all-gather of the partial roots, top levels on every rank ----

}  // extern "C"

namespace {

// How a tree of world * n_local leaves splits over `world` contiguous slices: each rank reduces
// `local_levels` levels (while its slice is whole sub-trees), leaving `roots_per_rank` partial roots; the gathered
// world * roots_per_rank partial roots must form a complete tree of `top_levels` levels.
struct ShardPlan {
    int local_levels = 0;
    size_t roots_per_rank = 0;
    int top_levels = 0;
};

int shard_plan(int arity, size_t n_local, int world, ShardPlan* plan) {
    if (n_local == 0 || world < 1) return ANEMOI_B200_ERR_LENGTH;
    size_t m = n_local;
    int local_levels = 0;
    while (m > 1 && m % (size_t)arity == 0) {
        m /= (size_t)arity;
        local_levels++;
    }
    size_t t = m * (size_t)world;
    int top_levels = 0;
    while (t > 1) {
        if (t % (size_t)arity) return ANEMOI_B200_ERR_LENGTH;  // the ranks do not split this tree into whole sub-trees
        t /= (size_t)arity;
        top_levels++;
    }
    plan->local_levels = local_levels;
    plan->roots_per_rank = m;
    plan->top_levels = top_levels;
    return ANEMOI_B200_OK;
}

int nccl_fail(int code, const char* what) {
    const anemoi::nccl::Api& nc = anemoi::nccl::api();
    snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", what, nc.ok ? nc.GetErrorString(code) : "libnccl.so.2 could not be loaded");
    return ANEMOI_B200_ERR_NCCL;
}
#define NK(call)                                        \
    do {                                                \
        int r__ = (call);                               \
        if (r__ != anemoi::nccl::kSuccess) return nccl_fail(r__, #call); \
    } while (0)

// Single-process communicators over devices 0..n-1 (ncclCommInitAll), created once per n and kept for the life of
// the process (communicator setup costs ~100 ms; the all-gather itself is microseconds).
std::mutex g_comm_mu;
std::mutex g_multi_gpu_turn;
std::vector<anemoi::nccl::comm_t> g_all_comms[kMaxDevices + 1];

int single_process_comms(int n_gpus, std::vector<anemoi::nccl::comm_t>* out) {
    const anemoi::nccl::Api& nc = anemoi::nccl::api();
    if (!nc.ok) return nccl_fail(-1, "NCCL");
    std::lock_guard<std::mutex> lock(g_comm_mu);
    std::vector<anemoi::nccl::comm_t>& c = g_all_comms[n_gpus];
    if (c.empty()) {
        std::vector<int> devs(n_gpus);
        for (int g = 0; g < n_gpus; g++) devs[g] = g;
        std::vector<anemoi::nccl::comm_t> fresh(n_gpus, nullptr);
        NK(nc.CommInitAll(fresh.data(), n_gpus, devs.data()));
        c = fresh;
    }
    *out = c;
    return ANEMOI_B200_OK;
}

}  // namespace

extern "C" {

int anemoi_b200_nccl_version(void) {
    const anemoi::nccl::Api& nc = anemoi::nccl::api();
    int v = 0;
    if (!nc.ok || !nc.GetVersion || nc.GetVersion(&v) != anemoi::nccl::kSuccess) return 0;
    return v;
}

int anemoi_b200_comm_unique_id(uint8_t* id128) {
    if (!id128) return ANEMOI_B200_ERR_ARG;
    const anemoi::nccl::Api& nc = anemoi::nccl::api();
    if (!nc.ok) return nccl_fail(-1, "NCCL");
    anemoi::nccl::unique_id id;
    NK(nc.GetUniqueId(&id));
    memcpy(id128, id.internal, sizeof(id.internal));
    return ANEMOI_B200_OK;
}

int anemoi_b200_comm_init_rank(const uint8_t* id128, int nranks, int rank, void** comm) {
    if (!id128 || !comm || nranks < 1 || rank < 0 || rank >= nranks) return ANEMOI_B200_ERR_ARG;
    const anemoi::nccl::Api& nc = anemoi::nccl::api();
    if (!nc.ok) return nccl_fail(-1, "NCCL");
    anemoi::nccl::unique_id id;
    memcpy(id.internal, id128, sizeof(id.internal));
    anemoi::nccl::comm_t c = nullptr;
    NK(nc.CommInitRank(&c, nranks, id, rank));  // collective over the nranks callers; binds the CURRENT device
    *comm = c;
    return ANEMOI_B200_OK;
}

int anemoi_b200_comm_destroy(void* comm) {
    if (!comm) return ANEMOI_B200_OK;
    const anemoi::nccl::Api& nc = anemoi::nccl::api();
    if (!nc.ok) return nccl_fail(-1, "NCCL");
    NK(nc.CommDestroy((anemoi::nccl::comm_t)comm));
    return ANEMOI_B200_OK;
}

int anemoi_b200_comm_info(void* comm, int* nranks, int* rank) {
    if (!comm) return ANEMOI_B200_ERR_ARG;
    const anemoi::nccl::Api& nc = anemoi::nccl::api();
    if (!nc.ok) return nccl_fail(-1, "NCCL");
    if (nranks) NK(nc.CommCount((anemoi::nccl::comm_t)comm, nranks));
    if (rank) NK(nc.CommUserRank((anemoi::nccl::comm_t)comm, rank));
    return ANEMOI_B200_OK;
}

size_t anemoi_b200_merkle_sharded_scratch_felts(int arity, size_t n_local, int nranks) {
    ShardPlan plan;
    if (arity < 2 || nranks < 1 || shard_plan(arity, n_local, nranks, &plan) != ANEMOI_B200_OK) return 0;
    const size_t gathered = plan.roots_per_rank * (size_t)nranks;
    return anemoi_b200_merkle_scratch_felts(arity, n_local) + plan.roots_per_rank + gathered +
           anemoi_b200_merkle_scratch_felts(arity, gathered);
}

int anemoi_b200_merkle_root_sharded_dev(int field, int inst, int arity, const uint64_t* d_local_leaves, size_t n_local,
                                        void* nccl_comm, uint64_t* d_scratch, uint64_t* d_root, void* stream) {
    int rc = merkle_check(field, inst, arity);
    if (rc) return rc;
    if (n_local == 0) return ANEMOI_B200_ERR_LENGTH;
    if (!d_local_leaves || !d_root) return ANEMOI_B200_ERR_ARG;
    int world = 1;
    if (nccl_comm) {  // NCCL is only bound (dlopen) when a communicator is actually in play
        const anemoi::nccl::Api& nc = anemoi::nccl::api();
        if (!nc.ok) return nccl_fail(-1, "NCCL");
        NK(nc.CommCount((anemoi::nccl::comm_t)nccl_comm, &world));
    }
    ShardPlan plan;
    rc = shard_plan(arity, n_local, world, &plan);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t words = (size_t)kFieldLimbs[field];
    const size_t fb = felt_bytes(field);
    DevBuf own;  // stream-ordered scratch from the library pool when the caller passes none
    if (!d_scratch) {
        cudaError_t e = own.alloc_async(anemoi_b200_merkle_sharded_scratch_felts(arity, n_local, world) * fb, st);
        if (e != cudaSuccess) return cuda_fail(e, "scratch allocation");
        d_scratch = (uint64_t*)own.p;
    }
    uint64_t* local_scratch = d_scratch;
    uint64_t* partial = local_scratch + anemoi_b200_merkle_scratch_felts(arity, n_local) * words;
    uint64_t* gathered = partial + plan.roots_per_rank * words;
    uint64_t* top_scratch = gathered + plan.roots_per_rank * (size_t)world * words;
    if (world == 1)  // whole tree on this device: roots_per_rank == 1 by construction of the plan
        return anemoi_b200_merkle_reduce_dev(field, inst, arity, d_local_leaves, n_local, plan.local_levels + plan.top_levels,
                                             local_scratch, d_root, st);
    rc = anemoi_b200_merkle_reduce_dev(field, inst, arity, d_local_leaves, n_local, plan.local_levels, local_scratch, partial, st);
    if (rc) return rc;
    // the one exchange step of the path: <= 2 field elements (<= 96 bytes) per rank over NVLink
    const anemoi::nccl::Api& nc = anemoi::nccl::api();
    NK(nc.AllGather(partial, gathered, plan.roots_per_rank * fb, anemoi::nccl::kUint8, (anemoi::nccl::comm_t)nccl_comm, st));
    return anemoi_b200_merkle_reduce_dev(field, inst, arity, gathered, plan.roots_per_rank * (size_t)world, plan.top_levels,
                                         top_scratch, d_root, st);
}

int anemoi_b200_merkle_root(int field, int inst, int arity, const uint64_t* leaves, size_t n_leaves, uint64_t* root,
                            int n_gpus) {
    int rc = merkle_check(field, inst, arity);
    if (rc) return rc;
    int height = 0;
    rc = tree_height(arity, n_leaves, &height);  // n_leaves must be arity^h
    if (rc) return rc;
    if (!leaves || !root) return ANEMOI_B200_ERR_ARG;
    if (n_gpus < 1 || (n_gpus & (n_gpus - 1)) || n_gpus > kMaxDevices) return ANEMOI_B200_ERR_ARG;
    const int count = anemoi_b200_device_count();
    if (count == 0) {
        snprintf(g_cuda_err, sizeof(g_cuda_err), "no CUDA device");
        return ANEMOI_B200_ERR_NO_DEVICE;
    }
    if (n_gpus > count) return ANEMOI_B200_ERR_ARG;
    if ((size_t)n_gpus > n_leaves) n_gpus = 1;
    const size_t fb = felt_bytes(field);
    const size_t words = (size_t)kFieldLimbs[field];
    const size_t slice = n_leaves / (size_t)n_gpus;
    ShardPlan plan;
    rc = shard_plan(arity, slice, n_gpus, &plan);
    if (rc) return rc;
    std::vector<anemoi::nccl::comm_t> comms(n_gpus, nullptr);
    // the cached single-process communicators serve one collective at a time: concurrent multi-GPU calls take turns
    std::unique_lock<std::mutex> turn(g_multi_gpu_turn, std::defer_lock);
    if (n_gpus > 1) {
        turn.lock();
        rc = single_process_comms(n_gpus, &comms);
        if (rc) return rc;
    }
    // one host thread + stream per device. Phase 1: device buffers (library pool) + H2D of the slice. Phase 2 (only when
    // every device got through phase 1, so that no rank is missing from the collective): sub-tree, all-gather, top levels
    // -- redundantly on every device, as in the multi-process form; device 0 returns the root.
    struct PerDevice {
        cudaStream_t st = nullptr, copy = nullptr;
        DevBuf leaves, level1, root;
        int rc = ANEMOI_B200_OK;
        std::string err;
    };
    std::vector<PerDevice> dev(n_gpus);
    // Big slices are uploaded in chunks on a second stream while the FIRST tree level of the chunks already there is
    // being hashed (a level-1 node needs only its own `arity` leaves): the 2 GiB upload of a 2^26-leaf tree hides behind
    // compute instead of preceding it. The remaining levels then start from the level-1 array.
    const size_t level1_nodes = plan.local_levels >= 1 ? slice / (size_t)arity : 0;
    const int chunks = (level1_nodes >= ((size_t)1 << 20) && level1_nodes % 8 == 0) ? 8 : 0;
    auto run_phase = [&](int phase) {
        auto body = [&](int g) {
            PerDevice& d = dev[g];
            d.rc = [&]() -> int {
                DeviceScope scope(g);
                if (scope.rc != ANEMOI_B200_OK) return scope.rc;
                if (phase == 1) {
                    CK(cudaStreamCreateWithFlags(&d.st, cudaStreamNonBlocking));
                    cudaError_t e = d.leaves.alloc_async(slice * fb, d.st);
                    if (e == cudaSuccess) e = d.root.alloc_async(fb, d.st);
                    if (e == cudaSuccess && chunks) e = d.level1.alloc_async(level1_nodes * fb, d.st);
                    if (e != cudaSuccess) return cuda_fail(e, "device allocation");
                    const uint64_t* src = leaves + (size_t)g * slice * words;
                    if (!chunks) {
                        CK(cudaMemcpyAsync(d.leaves.p, src, slice * fb, cudaMemcpyHostToDevice, d.st));
                        return ANEMOI_B200_OK;
                    }
                    CK(cudaStreamCreateWithFlags(&d.copy, cudaStreamNonBlocking));
                    CK(wait_for(d.copy, d.st));  // the buffers exist (stream-ordered allocation) before the copy stream writes into them
                    const size_t nodes_per_chunk = level1_nodes / (size_t)chunks, leaves_per_chunk = nodes_per_chunk * (size_t)arity;
                    const int mode = arity == 2 ? anemoi::MODE_COMPRESS : anemoi::MODE_COMPRESS4;
                    for (int c = 0; c < chunks; c++) {
                        uint64_t* d_chunk = (uint64_t*)d.leaves.p + (size_t)c * leaves_per_chunk * words;
                        CK(cudaMemcpyAsync(d_chunk, src + (size_t)c * leaves_per_chunk * words, leaves_per_chunk * fb,
                                           cudaMemcpyHostToDevice, d.copy));
                        CK(wait_for(d.st, d.copy));
                        int r1 = launch(field, inst, mode, d_chunk, (uint64_t*)d.level1.p + (size_t)c * nodes_per_chunk * words, nullptr,
                                        nodes_per_chunk, 0, d.st);
                        if (r1) return r1;
                    }
                    return ANEMOI_B200_OK;
                }
                // the sharded entry continues from the level-1 array when phase 1 already hashed the leaves
                const uint64_t* d_from = chunks ? (const uint64_t*)d.level1.p : (const uint64_t*)d.leaves.p;
                int r = anemoi_b200_merkle_root_sharded_dev(field, inst, arity, d_from, chunks ? level1_nodes : slice, comms[g], nullptr,
                                                            (uint64_t*)d.root.p, d.st);
                cudaError_t e = cudaSuccess;
                if (!r && g == 0 && (e = cudaMemcpyAsync(root, d.root.p, fb, cudaMemcpyDeviceToHost, d.st)) != cudaSuccess)
                    r = cuda_fail(e, "D2H root");
                e = cudaStreamSynchronize(d.st);
                if (!r && e != cudaSuccess) r = cuda_fail(e, "cudaStreamSynchronize");
                return r;
            }();
            d.err = g_cuda_err;
        };
        if (n_gpus == 1) {
            body(0);
        } else {
            std::vector<std::thread> th;
            for (int g = 0; g < n_gpus; g++) th.emplace_back(body, g);
            for (auto& t : th) t.join();
        }
        for (int g = 0; g < n_gpus; g++)
            if (dev[g].rc) {
                snprintf(g_cuda_err, sizeof(g_cuda_err), "gpu %d: %s", g, dev[g].err.c_str());
                return dev[g].rc;
            }
        return (int)ANEMOI_B200_OK;
    };
    rc = run_phase(1);
    if (!rc) rc = run_phase(2);
    for (int g = 0; g < n_gpus; g++) {  // stream-ordered frees, then the streams
        PerDevice& d = dev[g];
        if (!d.st) continue;
        DeviceScope scope(g);
        d.leaves.release();
        d.level1.release();
        d.root.release();
        cudaStreamSynchronize(d.st);
        if (d.copy) {
            cudaStreamSynchronize(d.copy);
            cudaStreamDestroy(d.copy);
        }
        cudaStreamDestroy(d.st);
    }
    return rc;
}

int anemoi_b200_merkle_open(int field, int inst, int arity, const uint64_t* leaves, size_t n_leaves,
                            const uint64_t* indices, size_t n_idx, uint64_t* root, uint64_t* paths, int device) {
    int rc = merkle_check(field, inst, arity);
    if (rc) return rc;
    int height = 0;
    rc = tree_height(arity, n_leaves, &height);
    if (rc) return rc;
    if (!leaves || !root || (n_idx && (!indices || (height && !paths)))) return ANEMOI_B200_ERR_ARG;
    for (size_t i = 0; i < n_idx; i++)
        if (indices[i] >= n_leaves) return ANEMOI_B200_ERR_ARG;
    const size_t fb = felt_bytes(field);
    if (height == 0) {
        memcpy(root, leaves, fb);
        return ANEMOI_B200_OK;
    }
    DeviceScope scope(device);
    if (scope.rc != ANEMOI_B200_OK) return scope.rc;
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    {
        const size_t tree_felts = anemoi_b200_merkle_tree_felts(arity, n_leaves);
        const size_t path_felts = n_idx * (size_t)height * (size_t)(arity - 1);
        DevBuf d_leaves, d_tree, d_idx, d_paths;
        cudaError_t e = d_leaves.alloc_async(n_leaves * fb, st);
        if (e == cudaSuccess) e = d_tree.alloc_async(tree_felts * fb, st);
        if (e == cudaSuccess) e = d_idx.alloc_async(n_idx * sizeof(uint64_t), st);
        if (e == cudaSuccess) e = d_paths.alloc_async(path_felts * fb, st);
        if (e != cudaSuccess) rc = cuda_fail(e, "device allocation");
        if (!rc && (e = cudaMemcpyAsync(d_leaves.p, leaves, n_leaves * fb, cudaMemcpyHostToDevice, st)) != cudaSuccess)
            rc = cuda_fail(e, "H2D leaves");
        if (!rc && n_idx && (e = cudaMemcpyAsync(d_idx.p, indices, n_idx * sizeof(uint64_t), cudaMemcpyHostToDevice, st)) != cudaSuccess)
            rc = cuda_fail(e, "H2D indices");
        if (!rc) rc = anemoi_b200_merkle_tree_dev(field, inst, arity, (const uint64_t*)d_leaves.p, n_leaves, (uint64_t*)d_tree.p, st);
        if (!rc)
            rc = anemoi_b200_merkle_open_dev(field, inst, arity, (const uint64_t*)d_leaves.p, (const uint64_t*)d_tree.p, n_leaves,
                                             (const uint64_t*)d_idx.p, n_idx, (uint64_t*)d_paths.p, st);
        if (!rc && (e = cudaMemcpyAsync(root, (const uint8_t*)d_tree.p + (tree_felts - 1) * fb, fb, cudaMemcpyDeviceToHost, st)) != cudaSuccess)
            rc = cuda_fail(e, "D2H root");
        if (!rc && path_felts && (e = cudaMemcpyAsync(paths, d_paths.p, path_felts * fb, cudaMemcpyDeviceToHost, st)) != cudaSuccess)
            rc = cuda_fail(e, "D2H paths");
        e = cudaStreamSynchronize(st);
        if (!rc && e != cudaSuccess) rc = cuda_fail(e, "cudaStreamSynchronize");
    }
    cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    return rc;
}

int anemoi_b200_merkle_verify(int field, int inst, int arity, const uint64_t* leaf_values, const uint64_t* indices,
                              const uint64_t* paths, int height, size_t n_idx, uint64_t* roots, int device) {
    int rc = merkle_check(field, inst, arity);
    if (rc) return rc;
    if (height < 0) return ANEMOI_B200_ERR_ARG;
    if (n_idx == 0) return ANEMOI_B200_OK;
    if (!leaf_values || !roots || (height && (!indices || !paths))) return ANEMOI_B200_ERR_ARG;
    const size_t fb = felt_bytes(field);
    if (height == 0) {
        memcpy(roots, leaf_values, n_idx * fb);
        return ANEMOI_B200_OK;
    }
    DeviceScope scope(device);
    if (scope.rc != ANEMOI_B200_OK) return scope.rc;
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    {
        const size_t path_felts = n_idx * (size_t)height * (size_t)(arity - 1);
        DevBuf d_vals, d_idx, d_paths, d_scratch, d_roots;
        cudaError_t e = d_vals.alloc_async(n_idx * fb, st);
        if (e == cudaSuccess) e = d_idx.alloc_async(n_idx * sizeof(uint64_t), st);
        if (e == cudaSuccess) e = d_paths.alloc_async(path_felts * fb, st);
        if (e == cudaSuccess) e = d_scratch.alloc_async(n_idx * (size_t)(arity + 1) * fb, st);
        if (e == cudaSuccess) e = d_roots.alloc_async(n_idx * fb, st);
        if (e != cudaSuccess) rc = cuda_fail(e, "device allocation");
        if (!rc && (e = cudaMemcpyAsync(d_vals.p, leaf_values, n_idx * fb, cudaMemcpyHostToDevice, st)) != cudaSuccess) rc = cuda_fail(e, "H2D");
        if (!rc && (e = cudaMemcpyAsync(d_idx.p, indices, n_idx * sizeof(uint64_t), cudaMemcpyHostToDevice, st)) != cudaSuccess) rc = cuda_fail(e, "H2D");
        if (!rc && (e = cudaMemcpyAsync(d_paths.p, paths, path_felts * fb, cudaMemcpyHostToDevice, st)) != cudaSuccess) rc = cuda_fail(e, "H2D");
        if (!rc)
            rc = anemoi_b200_merkle_verify_dev(field, inst, arity, (const uint64_t*)d_vals.p, (const uint64_t*)d_idx.p,
                                               (const uint64_t*)d_paths.p, height, n_idx, (uint64_t*)d_scratch.p, (uint64_t*)d_roots.p, st);
        if (!rc && (e = cudaMemcpyAsync(roots, d_roots.p, n_idx * fb, cudaMemcpyDeviceToHost, st)) != cudaSuccess) rc = cuda_fail(e, "D2H");
        e = cudaStreamSynchronize(st);
        if (!rc && e != cudaSuccess) rc = cuda_fail(e, "cudaStreamSynchronize");
    }
    cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    return rc;
}

}  // extern "C"
