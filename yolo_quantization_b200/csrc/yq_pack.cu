// Packed-weight arena: the kernel-layout filter images the prepare steps build from a layer's OIHW
// `weights_uint8` (parser.c:1124-1159 stream order) -- OHWI rows padded to the channel stride for the flat kernels,
// Toeplitz / even-odd tiles for the rows kernel, im2col-row tiles for the small-c kernel -- kept under a content key,
// so that a later load of the same weights uploads them without repacking.
//
//   key   = FNV-1a over (layout version, layer shape, per-channel zero points, the u8 weights) + a tag naming the image
//           and its flavour parameters
//   file  = "YQPK" u32 version u32 count, then per entry: u64 key, char tag[24], u64 bytes, u64 FNV-1a of the data, data
//
// The weights themselves are part of the key, so a stale or foreign file misses; an entry whose stored checksum does not match
// its data (corruption, truncation-and-rewrite) is dropped at load and counts as a miss.  What the format does NOT defend against
// is a 64-bit FNV collision between two different weight sets or deliberate tampering that keeps key and checksum consistent: the
// arena is a cache next to the .weights file it was built from, not an authenticated container.
// The in-memory arena only collects images while it is ENABLED (yq_pack_arena_enable / a successful yq_pack_arena_load): a
// process that never asks for an arena keeps no second copy of its filter images.  All entry points take one mutex.
#include <stdio.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "yq_common.h"

namespace {

constexpr uint32_t PACK_MAGIC = 0x4B505159u;   // "YQPK"
constexpr uint32_t PACK_VERSION = 5;           // bump when any kernel's filter image or the file format changes

struct Arena {
    std::map<std::pair<uint64_t, std::string>, std::vector<uint8_t>> entries;
    int hits = 0, misses = 0;
    bool dirty = false, enabled = false;
};
Arena g_arena;
std::mutex g_mu;

inline uint64_t fnv(uint64_t h, const void *p, size_t n)
{
    const uint8_t *b = (const uint8_t *)p;
    for (size_t i = 0; i < n; ++i) h = (h ^ b[i]) * 1099511628211ull;
    return h;
}

}  // namespace

namespace yq {

uint64_t pack_layer_key(const yq_conv_layer *l)
{
    const int shape[] = {(int)PACK_VERSION, l->h, l->w, l->c, l->cs_in, l->n, l->cs_out, l->size, l->stride, l->pad, l->zp_in};
    uint64_t h = fnv(14695981039346656037ull, shape, sizeof shape);
    h = fnv(h, l->host_zw.data(), l->host_zw.size() * sizeof(l->host_zw[0]));      // the rows kernel's signed blocks are w - zp_w
    return fnv(h, l->host_w.data(), l->host_w.size());
}

bool pack_fetch(const yq_conv_layer *l, const char *tag, std::vector<uint8_t> &out)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_arena.enabled) return false;
    auto it = g_arena.entries.find(std::make_pair(l->pack_key, std::string(tag)));
    if (it == g_arena.entries.end()) {
        ++g_arena.misses;
        return false;
    }
    ++g_arena.hits;
    out = it->second;
    return true;
}

void pack_put(const yq_conv_layer *l, const char *tag, const std::vector<uint8_t> &img)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_arena.enabled) return;
    g_arena.entries[std::make_pair(l->pack_key, std::string(tag))] = img;
    g_arena.dirty = true;
}


// ---- the arena as one blob (the file format, in memory) and its device-resident form for data-parallel replicas (yq_dp.cu) ----
void pack_serialize(std::vector<uint8_t> &blob, std::vector<PackIndexEntry> &index)
{
    std::lock_guard<std::mutex> lk(g_mu);
    blob.clear();
    index.clear();
    auto put = [&](const void *p, size_t n) { blob.insert(blob.end(), (const uint8_t *)p, (const uint8_t *)p + n); };
    const uint32_t hdr[3] = {PACK_MAGIC, PACK_VERSION, (uint32_t)g_arena.entries.size()};
    put(hdr, sizeof hdr);
    for (const auto &e : g_arena.entries) {
        char tag[24] = {0};
        strncpy(tag, e.first.second.c_str(), sizeof tag - 1);
        const uint64_t key = e.first.first, bytes = e.second.size(), sum = fnv(14695981039346656037ull, e.second.data(), e.second.size());
        put(&key, 8);
        put(tag, sizeof tag);
        put(&bytes, 8);
        put(&sum, 8);
        while (blob.size() % 256) blob.push_back(0);          // images start 256-byte aligned inside the blob (device copies)
        index.push_back(PackIndexEntry{key, std::string(tag), blob.size(), (size_t)bytes});
        put(e.second.data(), e.second.size());
    }
}

namespace {
struct DeviceArena {
    const uint8_t *blob = nullptr;                 // on the thread's current device
    const std::vector<PackIndexEntry> *index = nullptr;
    int hits = 0;
};
thread_local DeviceArena tls_dev_arena;
}  // namespace

void pack_set_device_arena(const uint8_t *blob_dev, const std::vector<PackIndexEntry> *index)
{
    tls_dev_arena.blob = blob_dev;
    tls_dev_arena.index = index;
    tls_dev_arena.hits = 0;
}
int pack_device_arena_hits() { return tls_dev_arena.hits; }

bool pack_fetch_device(const yq_conv_layer *l, const char *tag, size_t bytes, void **dev_out)
{
    if (!tls_dev_arena.blob || !tls_dev_arena.index) return false;
    for (const auto &e : *tls_dev_arena.index)
        if (e.key == l->pack_key && e.bytes == bytes && e.tag == tag) {
            void *p = nullptr;
            if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) return false;
            if (cudaMemcpy(p, tls_dev_arena.blob + e.offset, bytes, cudaMemcpyDeviceToDevice) != cudaSuccess) {
                cudaFree(p);
                return false;
            }
            *dev_out = p;
            ++tls_dev_arena.hits;
            return true;
        }
    return false;
}

}  // namespace yq

extern "C" int yq_pack_arena_clear(void)
{
    std::lock_guard<std::mutex> lk(g_mu);
    g_arena = Arena();
    return 0;
}

extern "C" int yq_pack_arena_enable(int enable)
{
    std::lock_guard<std::mutex> lk(g_mu);
    g_arena.enabled = enable != 0;
    return 0;
}

extern "C" int yq_pack_arena_load(const char *path)
{
    FILE *f = path ? fopen(path, "rb") : nullptr;
    if (!f) return yq::fail("yq_pack_arena_load: cannot open %s", path ? path : "(null)");
    uint32_t hdr[3];
    if (fread(hdr, 4, 3, f) != 3 || hdr[0] != PACK_MAGIC) {
        fclose(f);
        return yq::fail("yq_pack_arena_load: %s is not a packed-weight arena", path);
    }
    std::lock_guard<std::mutex> lk(g_mu);
    g_arena.enabled = true;
    if (hdr[1] != PACK_VERSION) {      // another layout generation: every key would miss anyway
        fclose(f);
        return 0;
    }
    int n = 0, kept = 0;
    for (uint32_t i = 0; i < hdr[2]; ++i) {
        uint64_t key, bytes, sum;
        char tag[24];
        if (fread(&key, 8, 1, f) != 1 || fread(tag, 1, sizeof tag, f) != sizeof tag || fread(&bytes, 8, 1, f) != 1 || fread(&sum, 8, 1, f) != 1 ||
            bytes > (1ull << 32))
            break;
        tag[sizeof tag - 1] = 0;
        std::vector<uint8_t> data((size_t)bytes);
        if (bytes && fread(data.data(), 1, (size_t)bytes, f) != bytes) break;
        ++n;
        if (fnv(14695981039346656037ull, data.data(), data.size()) != sum) continue;   // damaged entry: a miss, the image is rebuilt
        g_arena.entries[std::make_pair(key, std::string(tag))] = std::move(data);
        ++kept;
    }
    fclose(f);
    if ((uint32_t)n != hdr[2]) return yq::fail("yq_pack_arena_load: %s is truncated (%d of %u entries)", path, n, hdr[2]);
    if (kept != n) g_arena.dirty = true;    // rewrite the file without the damaged entries
    return kept;
}

extern "C" int yq_pack_arena_save(const char *path)
{
    FILE *f = path ? fopen(path, "wb") : nullptr;
    if (!f) return yq::fail("yq_pack_arena_save: cannot create %s", path ? path : "(null)");
    std::lock_guard<std::mutex> lk(g_mu);
    const uint32_t hdr[3] = {PACK_MAGIC, PACK_VERSION, (uint32_t)g_arena.entries.size()};
    bool ok = fwrite(hdr, 4, 3, f) == 3;
    for (const auto &e : g_arena.entries) {
        char tag[24] = {0};
        strncpy(tag, e.first.second.c_str(), sizeof tag - 1);
        const uint64_t key = e.first.first, bytes = e.second.size(), sum = fnv(14695981039346656037ull, e.second.data(), e.second.size());
        ok = ok && fwrite(&key, 8, 1, f) == 1 && fwrite(tag, 1, sizeof tag, f) == sizeof tag && fwrite(&bytes, 8, 1, f) == 1 && fwrite(&sum, 8, 1, f) == 1 &&
             (bytes == 0 || fwrite(e.second.data(), 1, (size_t)bytes, f) == bytes);
    }
    ok = (fclose(f) == 0) && ok;
    if (!ok) return yq::fail("yq_pack_arena_save: short write to %s", path);
    g_arena.dirty = false;
    return (int)g_arena.entries.size();
}

extern "C" int yq_pack_arena_stats(int *entries, int *hits, int *misses, int *dirty)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (entries) *entries = (int)g_arena.entries.size();
    if (hits) *hits = g_arena.hits;
    if (misses) *misses = g_arena.misses;
    if (dirty) *dirty = g_arena.dirty ? 1 : 0;
    return 0;
}
