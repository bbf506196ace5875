// yq_dp.cu -- C-level data-parallel inference over the GPUs of one box, no Python / torch involved.
//
// The path shards by image: every GPU holds a full replica of the network, there is NO per-step collective.  The reference's
// multi-GPU pattern is one host thread per device that calls cuda_set_device first (src/network.c:930-937, train_networks
// :1164-1194; weights averaged through the host, no NCCL); here:
//   * replica 0 is loaded on devices[0] by the calling thread: cfg parse, .weights read, host prep and the packing of every
//     kernel-layout filter image -- with the packed-weight arena collecting those images;
//   * the arena is serialised into one blob (the .yqpk layout, images 256-byte aligned), uploaded to devices[0] and handed to
//     every other device by ONE ncclBroadcast (the "trivial NCCL broadcast of weights only" of the north star);
//   * one host thread per remaining device loads its replica: cfg / .weights are parsed again (small host work: scales, zero
//     points, multipliers), but every big filter image comes out of the blob on that device by a device-to-device copy -- no
//     packing and no host-to-device upload of weights on the replicas.
// NCCL is bound at run time (dlopen libnccl.so.2) so that the library has no link-time dependency on it and never drags a
// second NCCL into a process that already carries one (torch); a box without NCCL gets a clear error from yq_dp_load_network
// and keeps every single-GPU entry point.
// Steady state: yq_dp_network_predict_u8 enqueues H2D + forward on every replica's streams from the calling thread (the
// per-replica entry points are asynchronous) and then collects -- the GPUs run concurrently without a host thread each.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <string>
#include <thread>
#include <vector>

#include "yq_common.h"

namespace {

struct Nccl {
    void *so = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string &err)
    {
        if (so) return true;
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            so = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (so) break;
        }
        if (!so) {
            err = std::string("cannot dlopen libnccl.so.2: ") + dlerror();
            return false;
        }
#define YQ_SYM(field, sym)                                            \
    field = (decltype(field))dlsym(so, sym);                          \
    if (!field) {                                                     \
        err = std::string("libnccl lacks ") + sym;                    \
        return false;                                                 \
    }
        YQ_SYM(CommInitAll, "ncclCommInitAll");
        YQ_SYM(CommDestroy, "ncclCommDestroy");
        YQ_SYM(Broadcast, "ncclBroadcast");
        YQ_SYM(GroupStart, "ncclGroupStart");
        YQ_SYM(GroupEnd, "ncclGroupEnd");
        YQ_SYM(GetErrorString, "ncclGetErrorString");
#undef YQ_SYM
        return true;
    }
};
Nccl g_nccl;

}  // namespace

struct yq_dp_network {
    std::vector<int> devices;
    std::vector<yq_network *> nets;
    std::vector<int> slots;
    size_t in_bytes = 0, out_floats = 0;
    size_t blob_bytes = 0;
    int images_from_blob = 0;        // filter images the replicas 1.. took from the broadcast blob (device-to-device)
};

extern "C" void yq_dp_free_network(yq_dp_network *dp)
{
    if (!dp) return;
    for (auto *n : dp->nets) yq_free_network(n);
    delete dp;
}

extern "C" yq_dp_network *yq_dp_load_network(const char *cfg, const char *weights, int batch_per_device, const int *devices, int n_devices)
{
    yq::clear_error();
    if (!cfg || !weights || !devices || n_devices <= 0 || n_devices > 64) {
        yq::fail("yq_dp_load_network: bad argument");
        return nullptr;
    }
    if (yq_device_count() < n_devices) {
        yq::fail("yq_dp_load_network: %d devices requested, %d visible", n_devices, yq_device_count());
        return nullptr;
    }
    yq_dp_network *dp = new yq_dp_network();
    dp->devices.assign(devices, devices + n_devices);
    dp->nets.assign(n_devices, nullptr);
    dp->slots.assign(n_devices, -1);
    auto bail = [&](const char *why) -> yq_dp_network * {
        std::string keep = why ? why : yq_last_error();
        yq_dp_free_network(dp);
        yq_pack_arena_enable(0);
        yq::fail("yq_dp_load_network: %s", keep.c_str());
        return nullptr;
    };
    // ---- replica 0: parse, prepare, pack (the arena collects the images)
    yq_pack_arena_clear();
    yq_pack_arena_enable(1);
    dp->nets[0] = yq_load_network(cfg, weights, batch_per_device, devices[0]);
    if (!dp->nets[0]) return bail(nullptr);
    int c, h, w;
    yq_network_input_dims(dp->nets[0], &c, &h, &w);
    dp->in_bytes = (size_t)yq_network_batch(dp->nets[0]) * c * h * w;
    dp->out_floats = yq_network_output_floats(dp->nets[0]);
    if (n_devices == 1) {
        yq_pack_arena_enable(0);
        yq_pack_arena_clear();
        return dp;
    }
    // ---- the arena as one blob on devices[0], one ncclBroadcast to every other device
    std::vector<uint8_t> blob;
    std::vector<yq::PackIndexEntry> index;
    yq::pack_serialize(blob, index);
    yq_pack_arena_enable(0);
    yq_pack_arena_clear();                       // the host copy has done its job
    dp->blob_bytes = blob.size();
    std::string err;
    if (!g_nccl.load(err)) return bail(err.c_str());
    std::vector<ncclComm_t> comms(n_devices, nullptr);
    std::vector<uint8_t *> dblob(n_devices, nullptr);
    std::vector<cudaStream_t> streams(n_devices, nullptr);
    auto release = [&]() {
        for (int i = 0; i < n_devices; ++i) {
            cudaSetDevice(devices[i]);
            if (streams[i]) cudaStreamDestroy(streams[i]);
            cudaFree(dblob[i]);
            if (comms[i]) g_nccl.CommDestroy(comms[i]);
        }
        cudaSetDevice(devices[0]);
    };
    ncclResult_t nr = g_nccl.CommInitAll(comms.data(), n_devices, devices);
    if (nr != ncclSuccess) {
        err = std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(nr);
        return bail(err.c_str());
    }
    bool ok = true;
    for (int i = 0; i < n_devices && ok; ++i)
        ok = cudaSetDevice(devices[i]) == cudaSuccess && cudaMalloc((void **)&dblob[i], blob.size()) == cudaSuccess &&
             cudaStreamCreateWithFlags(&streams[i], cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaSetDevice(devices[0]) == cudaSuccess &&
         cudaMemcpyAsync(dblob[0], blob.data(), blob.size(), cudaMemcpyHostToDevice, streams[0]) == cudaSuccess;
    if (ok) {
        g_nccl.GroupStart();
        for (int i = 0; i < n_devices; ++i) {
            nr = g_nccl.Broadcast(dblob[0], dblob[i], blob.size(), ncclUint8, 0, comms[i], streams[i]);
            if (nr != ncclSuccess) ok = false;
        }
        nr = g_nccl.GroupEnd();
        if (nr != ncclSuccess) ok = false;
        for (int i = 0; i < n_devices; ++i) ok = cudaSetDevice(devices[i]) == cudaSuccess && cudaStreamSynchronize(streams[i]) == cudaSuccess && ok;
    }
    if (!ok) {
        err = std::string("arena broadcast failed: ") + (nr != ncclSuccess ? g_nccl.GetErrorString(nr) : cudaGetErrorString(cudaGetLastError()));
        release();
        return bail(err.c_str());
    }
    // ---- one host thread per remaining device (network.c:930-937: each sets its device first)
    std::vector<std::string> errors(n_devices);
    std::vector<int> hits(n_devices, 0);
    std::vector<std::thread> th;
    for (int i = 1; i < n_devices; ++i)
        th.emplace_back([&, i]() {
            if (cudaSetDevice(devices[i]) != cudaSuccess) {
                errors[i] = "cudaSetDevice failed";
                return;
            }
            yq::pack_set_device_arena(dblob[i], &index);
            dp->nets[i] = yq_load_network(cfg, weights, batch_per_device, devices[i]);
            if (!dp->nets[i]) errors[i] = yq_last_error();      // (thread-local message)
            hits[i] = yq::pack_device_arena_hits();
            yq::pack_set_device_arena(nullptr, nullptr);
        });
    for (auto &t : th) t.join();
    release();
    for (int i = 1; i < n_devices; ++i) {
        if (!dp->nets[i]) {
            err = "replica on device " + std::to_string(devices[i]) + ": " + errors[i];
            return bail(err.c_str());
        }
        dp->images_from_blob += hits[i];
    }
    return dp;
}

extern "C" int yq_dp_num_devices(const yq_dp_network *dp) { return dp ? (int)dp->nets.size() : 0; }
extern "C" yq_network *yq_dp_replica(yq_dp_network *dp, int i) { return dp && i >= 0 && i < (int)dp->nets.size() ? dp->nets[i] : nullptr; }
extern "C" size_t yq_dp_arena_bytes(const yq_dp_network *dp) { return dp ? dp->blob_bytes : 0; }
extern "C" int yq_dp_images_from_arena(const yq_dp_network *dp) { return dp ? dp->images_from_blob : 0; }

// the 2-deep pipelined form (yq_network_submit_u8 / yq_network_collect on every replica): H2D, forward and D2H of neighbouring
// steps overlap on every device.  submit returns the slot to hand to collect (the replicas advance in lockstep), < 0 on error.
extern "C" int yq_dp_network_submit_u8(yq_dp_network *dp, const uint8_t *in_host)
{
    if (!dp || !in_host) return yq::fail("yq_dp_network_submit_u8: null argument");
    int slot = -1;
    for (size_t i = 0; i < dp->nets.size(); ++i) {
        const int s = yq_network_submit_u8(dp->nets[i], in_host + i * dp->in_bytes);
        if (s < 0) return -1;
        if (i && s != slot) return yq::fail("yq_dp_network_submit_u8: the replicas' pipelines are out of step (mixing per-replica and yq_dp calls?)");
        slot = s;
    }
    return slot;
}
extern "C" int yq_dp_network_collect(yq_dp_network *dp, int slot, float *out_host)
{
    if (!dp || !out_host) return yq::fail("yq_dp_network_collect: null argument");
    int rc = 0;
    for (size_t i = 0; i < dp->nets.size(); ++i)
        if (yq_network_collect(dp->nets[i], slot, out_host + i * dp->out_floats)) rc = -1;
    return rc;
}

// network_predict for n_devices * batch images: in_host [n_devices][batch][c][h][w] uint8 (image block i goes to device i),
// out_host [n_devices][yq_network_output_floats].  Pinned host memory (yq_host_alloc) lets the copies of all devices overlap.
extern "C" int yq_dp_network_predict_u8(yq_dp_network *dp, const uint8_t *in_host, float *out_host)
{
    if (!dp || !in_host || !out_host) return yq::fail("yq_dp_network_predict_u8: null argument");
    const int n = (int)dp->nets.size();
    for (int i = 0; i < n; ++i) {
        dp->slots[i] = yq_network_submit_u8(dp->nets[i], in_host + (size_t)i * dp->in_bytes);
        if (dp->slots[i] < 0) return -1;
    }
    int rc = 0;
    for (int i = 0; i < n; ++i)
        if (yq_network_collect(dp->nets[i], dp->slots[i], out_host + (size_t)i * dp->out_floats)) rc = -1;
    return rc;
}
