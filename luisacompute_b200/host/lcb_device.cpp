// lcb_device.cpp — host-side device layer (C++), the counterpart of the reference's CUDA backend host classes for
// the ray-tracing path.  It owns no kernels: all device work goes through the thin kernel C-ABI (include/lcb200_rt.h).
//
// Class-by-class correspondence (paths relative to /root/reference):
//   Primitive / Mesh / ProceduralPrimitive  <- CUDAPrimitive / CUDAMesh / CUDAProceduralPrimitive
//        src/backends/cuda/cuda_primitive.cpp:20-132, cuda_mesh.cpp:31-69, cuda_procedural_primitive.cpp:26-55
//   Accel                                   <- CUDAAccel  src/backends/cuda/cuda_accel.cpp:152-318
//   Device                                  <- CUDADevice (create_mesh & co.)  src/backends/cuda/cuda_device.cpp:909-981
// Differences by design: no host stall for compaction (cuda_primitive.cpp:74, cuda_accel.cpp:90), scratch and blobs
// come from the stream-ordered pool (cudaMallocAsync) and are reused across rebuilds of the same size.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <memory>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "../../include/lcb200_device.h"
#include "../csrc/lcrt_common.h"

namespace lcb {

using lcrt::Modification;

#define LCB_CHECK(expr)                                                                                   \
    do {                                                                                                  \
        int rc_ = (expr);                                                                                 \
        if (rc_ != 0) {                                                                                   \
            fprintf(stderr, "[lcb200] %s failed: %s (%s) at %s:%d\n", #expr, lcrtGetErrorString(rc_),     \
                    lcrtGetLastCudaError(), __FILE__, __LINE__);                                          \
            return rc_;                                                                                   \
        }                                                                                                 \
    } while (0)
#define LCB_CUDA(expr)                                                                                    \
    do {                                                                                                  \
        cudaError_t e_ = (expr);                                                                          \
        if (e_ != cudaSuccess) {                                                                          \
            fprintf(stderr, "[lcb200] %s failed: %s at %s:%d\n", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return LCRT_ERROR_CUDA;                                                                       \
        }                                                                                                 \
    } while (0)

static uint32_t build_flags(const lcbAccelOption &o) {
    uint32_t f = 0;
    if (o.hint == 1u) f |= LCRT_BUILD_FAST_BUILD;
    if (o.allow_update) f |= LCRT_BUILD_ALLOW_UPDATE;
    if (o.allow_compaction) f |= LCRT_BUILD_ALLOW_COMPACTION;
    return f;
}

// What a resource needs from its device: stream-ordered release of device memory (no host stall on destroy), the
// "some BLAS moved" epoch and the list of structures whose device-side build status has not been looked at yet.
struct DeviceServices {
    virtual ~DeviceServices() = default;
    virtual void retire(void *p) = 0;      // free `p` once everything queued so far on the device's known streams is done
    virtual void note_stream(cudaStream_t s) = 0;
    std::atomic<uint64_t> blob_epoch{1};   // bumped whenever a primitive's blob changes address
};

struct Resource {
    enum class Tag { MESH, PROCEDURAL, ACCEL };
    Tag tag;
    lcbAccelOption option;
    DeviceServices *dev;
    std::mutex mutex;
    void *blob = nullptr;
    uint64_t blob_bytes = 0;
    void *temp = nullptr;// kept between builds of the same size (test_rtx rebuilds every frame, test_rtx.cpp:122-139)
    uint64_t temp_bytes = 0;
    int last_build_kind = 0;
    cudaEvent_t ready = nullptr;// recorded on the build stream after every build / import: consumers on other streams wait on it
    bool check_pending = false; // built since its header was last read back (lcbSynchronize surfaces build errors)
    Resource(Tag t, const lcbAccelOption &o, DeviceServices *d) : tag{t}, option{o}, dev{d} {}
    virtual ~Resource() {
        // stream-ordered: the memory goes back to the pool when the work queued so far has passed (the reference frees with
        // cuMemFree, an implicit device synchronisation)
        if (blob) dev->retire(blob);
        if (temp) dev->retire(temp);
        if (ready) cudaEventDestroy(ready);
    }
    uint64_t native() const { return reinterpret_cast<uint64_t>(blob); }

    int mark_ready(cudaStream_t stream) {
        if (!ready) LCB_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
        LCB_CUDA(cudaEventRecord(ready, stream));
        check_pending = true;
        return 0;
    }

    int ensure(cudaStream_t stream, uint64_t need_blob, uint64_t need_temp) {
        if (blob_bytes < need_blob || blob_bytes > need_blob * 2 + (1u << 20)) {
            if (blob) LCB_CUDA(cudaFreeAsync(blob, stream));
            blob = nullptr, blob_bytes = 0;
            LCB_CUDA(cudaMallocAsync(&blob, need_blob, stream));
            blob_bytes = need_blob;
            dev->blob_epoch.fetch_add(1);
        }
        if (temp_bytes < need_temp || temp_bytes > need_temp * 2 + (1u << 20)) {
            if (temp) LCB_CUDA(cudaFreeAsync(temp, stream));
            temp = nullptr, temp_bytes = 0;
            LCB_CUDA(cudaMallocAsync(&temp, need_temp, stream));
            temp_bytes = need_temp;
        }
        return 0;
    }
};

struct Primitive : Resource {
    using Resource::Resource;
    bool built = false;
    bool blob_updatable = false;// the blob carries the parent array a refit needs (imported blobs may not)
};

struct Mesh final : Primitive {
    uint64_t vertex_buffer = 0, vertex_buffer_size = 0, vertex_stride = 0, triangle_buffer = 0, triangle_buffer_size = 0;
    Mesh(const lcbAccelOption &o, DeviceServices *d) : Primitive{Tag::MESH, o, d} {}

    int build(cudaStream_t stream, const lcbMeshBuildCommand &c) {
        std::scoped_lock lock{mutex};
        const uint64_t vb = c.vertex_buffer + c.vertex_buffer_offset, tb = c.triangle_buffer + c.triangle_buffer_offset;
        if (c.vertex_stride < 12 || (c.vertex_stride % 4) != 0) {
            fprintf(stderr, "[lcb200] invalid vertex stride %llu\n", (unsigned long long)c.vertex_stride);
            return LCRT_ERROR_INVALID_VALUE;
        }
        // cuda_mesh.cpp:42-54; a blob without the parent array (imported from a build without allow_update) can not be refitted
        const bool requires_build = !built || !option.allow_update || !blob_updatable || c.request == LCB_BUILD_FORCE_BUILD || vb != vertex_buffer ||
                                    c.vertex_buffer_size != vertex_buffer_size || c.vertex_stride != vertex_stride || tb != triangle_buffer ||
                                    c.triangle_buffer_size != triangle_buffer_size;
        vertex_buffer = vb, vertex_buffer_size = c.vertex_buffer_size, vertex_stride = c.vertex_stride;
        triangle_buffer = tb, triangle_buffer_size = c.triangle_buffer_size;
        const uint32_t n = (uint32_t)(triangle_buffer_size / 12u);// sizeof(Triangle), cuda_mesh.cpp:25-26
        const uint32_t flags = build_flags(option);
        lcrtSizes sz;
        LCB_CHECK(lcrtAccelComputeMemoryUsage(LCRT_KIND_TRIANGLES, n, flags, &sz));
        if (requires_build) {
            LCB_CHECK(ensure(stream, sz.blob_bytes, sz.build_temp_bytes > sz.update_temp_bytes ? sz.build_temp_bytes : sz.update_temp_bytes));
            LCB_CHECK(lcrtBuildTriangles(stream, reinterpret_cast<const void *>(vb), vertex_stride, reinterpret_cast<const void *>(tb), n, flags, temp,
                                         temp_bytes, blob, blob_bytes));
            built = true, last_build_kind = 1, blob_updatable = option.allow_update != 0u;
        } else {
            LCB_CHECK(ensure(stream, blob_bytes, sz.update_temp_bytes));// an imported blob arrives without scratch
            LCB_CHECK(lcrtRefitTriangles(stream, reinterpret_cast<const void *>(vb), vertex_stride, reinterpret_cast<const void *>(tb), n, temp,
                                         temp_bytes, blob));
            last_build_kind = 2;
        }
        return mark_ready(stream);
    }
};

struct ProceduralPrimitive final : Primitive {
    uint64_t aabb_buffer = 0, aabb_buffer_size = 0;
    ProceduralPrimitive(const lcbAccelOption &o, DeviceServices *d) : Primitive{Tag::PROCEDURAL, o, d} {}

    int build(cudaStream_t stream, const lcbProceduralPrimitiveBuildCommand &c) {
        std::scoped_lock lock{mutex};
        const uint64_t ab = c.aabb_buffer + c.aabb_buffer_offset;
        // cuda_procedural_primitive.cpp:33-43
        const bool requires_build = !built || !option.allow_update || !blob_updatable || c.request == LCB_BUILD_FORCE_BUILD || ab != aabb_buffer ||
                                    c.aabb_buffer_size != aabb_buffer_size;
        aabb_buffer = ab, aabb_buffer_size = c.aabb_buffer_size;
        const uint32_t n = (uint32_t)(aabb_buffer_size / 24u);// sizeof(AABB), cuda_procedural_primitive.cpp:17-19
        const uint32_t flags = build_flags(option);
        lcrtSizes sz;
        LCB_CHECK(lcrtAccelComputeMemoryUsage(LCRT_KIND_AABBS, n, flags, &sz));
        if (requires_build) {
            LCB_CHECK(ensure(stream, sz.blob_bytes, sz.build_temp_bytes > sz.update_temp_bytes ? sz.build_temp_bytes : sz.update_temp_bytes));
            LCB_CHECK(lcrtBuildAabbs(stream, reinterpret_cast<const void *>(ab), n, flags, temp, temp_bytes, blob, blob_bytes));
            built = true, last_build_kind = 1, blob_updatable = option.allow_update != 0u;
        } else {
            LCB_CHECK(ensure(stream, blob_bytes, sz.update_temp_bytes));
            LCB_CHECK(lcrtRefitAabbs(stream, reinterpret_cast<const void *>(ab), n, temp, temp_bytes, blob));
            last_build_kind = 2;
        }
        return mark_ready(stream);
    }
};

static uint64_t next_pow2(uint64_t v) {
    uint64_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

using ResourceMap = std::unordered_map<uint64_t, std::shared_ptr<Resource>>;

struct Accel final : Resource {
    void *instance_buffer = nullptr;
    uint64_t instance_buffer_size = 0;// bytes
    uint32_t initialised_instances = 0;
    // per instance: the primitive it refers to and the blob address the instance record currently holds
    std::vector<Primitive *> primitives;
    std::vector<uint64_t> prim_handles;
    // the distinct primitives the instances refer to, kept alive (shared ownership) while they do: destroying a mesh that an
    // accel still instances can not leave a dangling pointer behind (the reference's CUDAAccel holds raw pointers)
    struct Held {
        std::shared_ptr<Primitive> prim;
        uint32_t uses = 0;
    };
    std::unordered_map<uint64_t, Held> held;      // lcb handle -> primitive
    std::unordered_map<Primitive *, uint64_t> held_key;
    uint64_t epoch_seen = 0;                       // DeviceServices::blob_epoch at the last handle scan
    bool requires_rebuild = true;
    bool built = false;
    // modification upload staging, reused across builds (pinned host + device); `staged` guards the pinned half
    Modification *h_mods = nullptr;
    void *d_mods = nullptr;
    uint32_t mods_capacity = 0;
    cudaEvent_t staged = nullptr;
    Accel(const lcbAccelOption &o, DeviceServices *d) : Resource{Tag::ACCEL, o, d} {}
    ~Accel() override {
        if (instance_buffer) dev->retire(instance_buffer);
        if (d_mods) dev->retire(d_mods);
        if (staged) {
            cudaEventSynchronize(staged);
            cudaEventDestroy(staged);
        }
        if (h_mods) cudaFreeHost(h_mods);
    }

    void release(Primitive *p) {
        auto k = held_key.find(p);
        if (k == held_key.end()) return;
        auto it = held.find(k->second);
        if (it != held.end() && --it->second.uses == 0u) {
            held.erase(it);
            held_key.erase(k);
        }
    }

    int stage_mods(cudaStream_t stream, uint32_t n) {
        if (n > mods_capacity) {
            if (staged) LCB_CUDA(cudaEventSynchronize(staged));
            if (h_mods) LCB_CUDA(cudaFreeHost(h_mods));
            if (d_mods) dev->retire(d_mods);
            h_mods = nullptr, d_mods = nullptr, mods_capacity = 0;
            const uint32_t cap = (uint32_t)next_pow2(n < 64u ? 64u : n);
            LCB_CUDA(cudaMallocHost(reinterpret_cast<void **>(&h_mods), (size_t)cap * sizeof(Modification)));
            LCB_CUDA(cudaMallocAsync(&d_mods, (size_t)cap * sizeof(Modification), stream));
            mods_capacity = cap;
        }
        if (!staged) LCB_CUDA(cudaEventCreateWithFlags(&staged, cudaEventDisableTiming));
        else LCB_CUDA(cudaEventSynchronize(staged));// the previous upload has left the pinned buffer (long ago, normally)
        return 0;
    }

    int build(cudaStream_t stream, const lcbAccelBuildCommand &c, const ResourceMap &resources) {
        std::scoped_lock lock{mutex};
        const uint32_t instance_count = c.instance_count;
        if (instance_count == 0u) {// cuda_accel.cpp:159
            fprintf(stderr, "[lcb200] Instance count must be greater than 0.\n");
            return LCRT_ERROR_INVALID_VALUE;
        }
        // ---- validate every modification before any state changes (an error leaves the accel as it was)
        const uint32_t n = c.modification_count;
        const Modification *in_mods = static_cast<const Modification *>(c.modifications);
        if (n && !in_mods) return LCRT_ERROR_INVALID_VALUE;
        std::unordered_map<uint64_t, std::shared_ptr<Primitive>> fresh;// primitives this batch introduces
        {
            uint64_t last_handle = ~0ull;
            for (uint32_t k = 0; k < n; k++) {
                const Modification &m = in_mods[k];
                if (m.index >= instance_count) {
                    fprintf(stderr, "[lcb200] modification index %u out of range\n", m.index);
                    return LCRT_ERROR_INVALID_VALUE;
                }
                if ((m.flags & lcrt::MOD_PRIMITIVE) && m.primitive != last_handle) {
                    last_handle = m.primitive;
                    if (held.count(m.primitive) || fresh.count(m.primitive)) continue;
                    auto it = resources.find(m.primitive);
                    if (it == resources.end() || it->second->tag == Resource::Tag::ACCEL) {
                        fprintf(stderr, "[lcb200] modification refers to an unknown primitive handle\n");
                        return LCRT_ERROR_INVALID_VALUE;
                    }
                    fresh.emplace(m.primitive, std::static_pointer_cast<Primitive>(it->second));
                }
            }
        }
        // ---- instance buffer growth to the next power of two, old contents preserved (cuda_accel.cpp:160-170)
        if (const uint64_t size = (uint64_t)instance_count * sizeof(lcrt::Instance); instance_buffer_size < size) {
            void *old = instance_buffer;
            const uint64_t new_size = next_pow2(size);
            void *grown = nullptr;
            LCB_CUDA(cudaMallocAsync(&grown, new_size, stream));
            if (old) {
                LCB_CUDA(cudaMemcpyAsync(grown, old, instance_buffer_size, cudaMemcpyDeviceToDevice, stream));
                LCB_CUDA(cudaFreeAsync(old, stream));
            }
            instance_buffer = grown, instance_buffer_size = new_size;
        }
        if (instance_count > initialised_instances) {
            // fresh slots start as identity / mask 0xff / opaque (accel.rs:284-295)
            LCB_CHECK(lcrtInstancesInit(stream, instance_buffer, initialised_instances, instance_count - initialised_instances));
            initialised_instances = instance_count;
        }
        const bool instance_count_changed = primitives.size() != instance_count;
        for (size_t i = instance_count; i < primitives.size(); i++)
            if (primitives[i]) release(primitives[i]);// popped instances let go of their primitive
        primitives.resize(instance_count, nullptr);
        prim_handles.resize(instance_count, 0ull);
        // ---- upload the modifications and apply them (cuda_accel.cpp:176-240)
        Modification few[16];// a handful of modifications go out in the launch parameters (lcrtInstancesApplyModsHost)
        const bool inline_mods = n != 0u && n <= 16u;
        if (n) {
            if (!inline_mods) LCB_CHECK(stage_mods(stream, n));
            Modification *h_mods = inline_mods ? few : this->h_mods;
            memcpy(h_mods, in_mods, (size_t)n * sizeof(Modification));
            uint64_t last_handle = ~0ull;
            Primitive *last_prim = nullptr;
            for (uint32_t k = 0; k < n; k++) {
                Modification &m = h_mods[k];
                if (!(m.flags & lcrt::MOD_PRIMITIVE)) continue;
                requires_rebuild = true;
                if (m.primitive != last_handle) {
                    last_handle = m.primitive;
                    auto h = held.find(m.primitive);
                    if (h == held.end()) {
                        h = held.emplace(m.primitive, Held{fresh.at(m.primitive), 0u}).first;
                        held_key.emplace(h->second.prim.get(), m.primitive);
                    }
                    last_prim = h->second.prim.get();
                }
                if (primitives[m.index] != last_prim) {
                    held.find(last_handle)->second.uses++;
                    if (primitives[m.index]) release(primitives[m.index]);
                    primitives[m.index] = last_prim;
                }
                m.primitive = last_prim->native();
                prim_handles[m.index] = last_prim->native();
                if (last_prim->tag == Resource::Tag::PROCEDURAL) m.flags |= lcrt::MOD_PROCEDURAL;
            }
            if (inline_mods) {
                LCB_CHECK(lcrtInstancesApplyModsHost(stream, instance_buffer, h_mods, n));
            } else {
                LCB_CUDA(cudaMemcpyAsync(d_mods, h_mods, (size_t)n * sizeof(Modification), cudaMemcpyHostToDevice, stream));
                LCB_CUDA(cudaEventRecord(staged, stream));
                LCB_CHECK(lcrtInstancesApplyMods(stream, instance_buffer, d_mods, n));
            }
        }
        // ---- BLAS handles that changed through a rebuild without a modification (cuda_accel.cpp:244-258, 286-308): the
        // reference scans every instance on every build; here the scan only runs when some primitive's blob has moved
        // since the last one (DeviceServices::blob_epoch)
        bool handles_changed = false;
        if (const uint64_t epoch = dev->blob_epoch.load(); epoch != epoch_seen) {
            epoch_seen = epoch;
            std::vector<lcrt::InstanceHandleUpdate> changed;
            for (uint32_t i = 0; i < instance_count; i++) {
                if (primitives[i] != nullptr && primitives[i]->native() != prim_handles[i]) {
                    changed.push_back({i, primitives[i]->native()});
                    prim_handles[i] = primitives[i]->native();
                }
            }
            if (!changed.empty()) {
                handles_changed = true;
                void *dbuf = nullptr;
                const size_t bytes = changed.size() * sizeof(lcrt::InstanceHandleUpdate);
                LCB_CUDA(cudaMallocAsync(&dbuf, bytes, stream));
                cudaError_t e = cudaMemcpyAsync(dbuf, changed.data(), bytes, cudaMemcpyHostToDevice, stream);// pageable: staged before it returns
                int rc = e == cudaSuccess ? lcrtInstancesUpdateHandles(stream, instance_buffer, dbuf, (uint32_t)changed.size()) : LCRT_ERROR_CUDA;
                cudaFreeAsync(dbuf, stream);
                if (rc != 0) return rc;
            }
        }
        // cuda_accel.cpp:261-267
        requires_rebuild = requires_rebuild || c.request == LCB_BUILD_FORCE_BUILD || !option.allow_update || !built || instance_count_changed ||
                           handles_changed;
        if (!c.update_instance_buffer_only) {// cuda_accel.cpp:309-317
            const uint32_t flags = build_flags(option);
            lcrtSizes sz;
            LCB_CHECK(lcrtAccelComputeMemoryUsage(LCRT_KIND_INSTANCES, instance_count, flags, &sz));
            if (requires_rebuild) {
                LCB_CHECK(ensure(stream, sz.blob_bytes, sz.build_temp_bytes > sz.update_temp_bytes ? sz.build_temp_bytes : sz.update_temp_bytes));
                LCB_CHECK(lcrtBuildTlas(stream, instance_buffer, instance_count, flags, temp, temp_bytes, blob, blob_bytes));
                built = true, last_build_kind = 1;
            } else {
                LCB_CHECK(lcrtRefitTlas(stream, instance_buffer, instance_count, temp, temp_bytes, blob));
                last_build_kind = 2;
            }
            requires_rebuild = false;
        }
        return mark_ready(stream);
    }
};

}// namespace lcb

struct lcbDevice final : lcb::DeviceServices {
    int index = 0;
    std::mutex mutex;
    lcb::ResourceMap resources;
    // streams this device has seen work on (builds, traces): a destroyed resource's memory is released behind all of them
    std::mutex stream_mutex;
    std::vector<cudaStream_t> streams;
    cudaStream_t reaper = nullptr;
    cudaEvent_t reaper_event = nullptr;
    // end-to-end helper state
    static constexpr int STAGE_SLOTS = 4;
    cudaStream_t io_streams[STAGE_SLOTS] = {};
    void *stage_rays[STAGE_SLOTS] = {};
    void *stage_hits[STAGE_SLOTS] = {};
    uint64_t stage_capacity = 0;

    void note_stream(cudaStream_t s) override {
        std::scoped_lock lock{stream_mutex};
        for (auto k : streams)
            if (k == s) return;
        streams.push_back(s);
    }
    void forget_stream(cudaStream_t s) {
        std::scoped_lock lock{stream_mutex};
        for (size_t i = 0; i < streams.size(); i++)
            if (streams[i] == s) {
                streams.erase(streams.begin() + (long)i);
                return;
            }
    }
    // the reaper stream waits for what is queued right now on every known stream, then frees: no host synchronisation
    void retire(void *p) override {
        if (!p) return;
        cudaSetDevice(index);
        std::scoped_lock lock{stream_mutex};
        bool ok = reaper != nullptr || cudaStreamCreateWithFlags(&reaper, cudaStreamNonBlocking) == cudaSuccess;
        ok = ok && (reaper_event != nullptr || cudaEventCreateWithFlags(&reaper_event, cudaEventDisableTiming) == cudaSuccess);
        for (size_t i = 0; ok && i < streams.size(); i++)
            ok = cudaEventRecord(reaper_event, streams[i]) == cudaSuccess && cudaStreamWaitEvent(reaper, reaper_event, 0) == cudaSuccess;
        if (ok && cudaFreeAsync(p, reaper) == cudaSuccess) return;
        // a stream handle that is no longer valid, or memory that did not come from the pool: drain and free the plain way
        cudaGetLastError();
        cudaDeviceSynchronize();
        if (cudaFreeAsync(p, nullptr) != cudaSuccess) {
            cudaGetLastError();
            cudaFree(p);
        }
    }

    template<class T>
    T *find(uint64_t handle, lcb::Resource::Tag tag) {
        std::scoped_lock lock{mutex};
        auto it = resources.find(handle);
        if (it == resources.end() || it->second->tag != tag) return nullptr;
        return static_cast<T *>(it->second.get());
    }
    lcb::Resource *find_any(uint64_t handle) {
        std::scoped_lock lock{mutex};
        auto it = resources.find(handle);
        return it == resources.end() ? nullptr : it->second.get();
    }
    template<class T>
    uint64_t add(const lcbAccelOption *option) {
        lcbAccelOption o{0u, 1u, 0u};// AccelOption defaults (resource.h:119-125)
        if (option) o = *option;
        auto r = std::make_shared<T>(o, this);
        const uint64_t handle = reinterpret_cast<uint64_t>(r.get());// the CUDA backend also uses object addresses
        std::scoped_lock lock{mutex};
        resources.emplace(handle, std::move(r));
        return handle;
    }
    void remove(uint64_t handle) {
        std::shared_ptr<lcb::Resource> victim;
        {
            std::scoped_lock lock{mutex};
            auto it = resources.find(handle);
            if (it == resources.end()) return;
            victim = std::move(it->second);
            resources.erase(it);
        }
        cudaSetDevice(index);
        victim.reset();// last owner (no accel instances it): ~Resource retires the device memory in stream order
    }
    // device-side build status of everything built since the last look (called after a stream synchronisation)
    int surface_build_errors() {
        std::vector<std::shared_ptr<lcb::Resource>> todo;
        {
            std::scoped_lock lock{mutex};
            for (auto &kv : resources)
                if (kv.second->check_pending) todo.push_back(kv.second);
        }
        int rc = 0;
        for (auto &r : todo) {
            std::scoped_lock lock{r->mutex};
            if (!r->check_pending || !r->blob || !r->ready) continue;
            if (cudaEventQuery(r->ready) != cudaSuccess) {
                cudaGetLastError();
                continue;// built on another stream that is still running: next time
            }
            r->check_pending = false;
            lcrt::BlobHeader h;
            if (cudaMemcpy(&h, r->blob, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) return LCRT_ERROR_CUDA;
            if (h.magic != lcrt::BLOB_MAGIC || h.error != 0u || !h.built) {
                fprintf(stderr, "[lcb200] acceleration structure %p: device-side build / refit error %u (built=%u)\n", (void *)r.get(), h.error,
                        h.built);
                rc = h.error == 4u ? LCRT_ERROR_NOT_UPDATABLE : LCRT_ERROR_BUILD_FAILED;
            }
            if (h.trace_error != 0u) {
                fprintf(stderr, "[lcb200] acceleration structure %p: a traversal ran out of stack (trace_error=%u)\n", (void *)r.get(), h.trace_error);
                rc = LCRT_ERROR_BUILD_FAILED;
            }
        }
        return rc;
    }
};

#define LCB_DEVICE_GUARD(device)                                             \
    if ((device) == nullptr) return LCRT_ERROR_INVALID_VALUE;                \
    if (cudaSetDevice((device)->index) != cudaSuccess) return LCRT_ERROR_CUDA

extern "C" {

const char *lcbBackendName(void) { return "b200rt"; }

int lcbDeviceCount(void) {
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

int lcbDeviceName(int device_index, char *name, int capacity) {
    cudaDeviceProp prop;
    if (name == nullptr || capacity <= 0 || cudaGetDeviceProperties(&prop, device_index) != cudaSuccess) return LCRT_ERROR_INVALID_VALUE;
    snprintf(name, (size_t)capacity, "%s", prop.name);
    return 0;
}

lcbDevice *lcbDeviceCreate(int device_index) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        fprintf(stderr, "[lcb200] no CUDA device available: this backend has no CPU fallback\n");
        return nullptr;
    }
    if (device_index < 0 || device_index >= count) {
        fprintf(stderr, "[lcb200] invalid device index %d (have %d)\n", device_index, count);
        return nullptr;
    }
    if (cudaSetDevice(device_index) != cudaSuccess) return nullptr;
    auto d = new lcbDevice;
    d->index = device_index;
    return d;
}

void lcbDeviceDestroy(lcbDevice *device) {
    if (!device) return;
    cudaSetDevice(device->index);
    cudaDeviceSynchronize();
    device->resources.clear();
    if (device->reaper) {
        cudaStreamSynchronize(device->reaper);
        cudaStreamDestroy(device->reaper);
    }
    if (device->reaper_event) cudaEventDestroy(device->reaper_event);
    for (int i = 0; i < lcbDevice::STAGE_SLOTS; i++) {
        if (device->stage_rays[i]) cudaFree(device->stage_rays[i]);
        if (device->stage_hits[i]) cudaFree(device->stage_hits[i]);
        if (device->io_streams[i]) cudaStreamDestroy(device->io_streams[i]);
    }
    delete device;
}

uint64_t lcbCreateMesh(lcbDevice *device, const lcbAccelOption *option) { return device ? device->add<lcb::Mesh>(option) : ~0ull; }
void lcbDestroyMesh(lcbDevice *device, uint64_t handle) {
    if (device) device->remove(handle);
}
uint64_t lcbCreateProceduralPrimitive(lcbDevice *device, const lcbAccelOption *option) {
    return device ? device->add<lcb::ProceduralPrimitive>(option) : ~0ull;
}
void lcbDestroyProceduralPrimitive(lcbDevice *device, uint64_t handle) {
    if (device) device->remove(handle);
}
uint64_t lcbCreateAccel(lcbDevice *device, const lcbAccelOption *option) { return device ? device->add<lcb::Accel>(option) : ~0ull; }
void lcbDestroyAccel(lcbDevice *device, uint64_t handle) {
    if (device) device->remove(handle);
}

int lcbMeshBuild(lcbDevice *device, lcrtStream stream, const lcbMeshBuildCommand *command) {
    LCB_DEVICE_GUARD(device);
    if (!command) return LCRT_ERROR_INVALID_VALUE;
    auto m = device->find<lcb::Mesh>(command->handle, lcb::Resource::Tag::MESH);
    if (!m) return LCRT_ERROR_INVALID_VALUE;
    device->note_stream(static_cast<cudaStream_t>(stream));
    return m->build(static_cast<cudaStream_t>(stream), *command);
}

int lcbProceduralPrimitiveBuild(lcbDevice *device, lcrtStream stream, const lcbProceduralPrimitiveBuildCommand *command) {
    LCB_DEVICE_GUARD(device);
    if (!command) return LCRT_ERROR_INVALID_VALUE;
    auto p = device->find<lcb::ProceduralPrimitive>(command->handle, lcb::Resource::Tag::PROCEDURAL);
    if (!p) return LCRT_ERROR_INVALID_VALUE;
    device->note_stream(static_cast<cudaStream_t>(stream));
    return p->build(static_cast<cudaStream_t>(stream), *command);
}

int lcbAccelBuild(lcbDevice *device, lcrtStream stream, const lcbAccelBuildCommand *command) {
    LCB_DEVICE_GUARD(device);
    if (!command) return LCRT_ERROR_INVALID_VALUE;
    auto a = device->find<lcb::Accel>(command->handle, lcb::Resource::Tag::ACCEL);
    if (!a) return LCRT_ERROR_INVALID_VALUE;
    device->note_stream(static_cast<cudaStream_t>(stream));
    std::scoped_lock lock{device->mutex};// the resource table must not change while the accel resolves primitive handles
    return a->build(static_cast<cudaStream_t>(stream), *command, device->resources);
}

int lcbAccelGetBinding(lcbDevice *device, uint64_t accel, lcbAccelBinding *out) {
    if (!device || !out) return LCRT_ERROR_INVALID_VALUE;
    auto a = device->find<lcb::Accel>(accel, lcb::Resource::Tag::ACCEL);
    if (!a) return LCRT_ERROR_INVALID_VALUE;
    out->handle = a->native(), out->instances = reinterpret_cast<uint64_t>(a->instance_buffer);
    return 0;
}

int lcbLastBuildKind(lcbDevice *device, uint64_t handle) {
    if (!device) return LCRT_ERROR_INVALID_VALUE;
    auto r = device->find_any(handle);
    return r ? r->last_build_kind : LCRT_ERROR_INVALID_VALUE;
}

uint64_t lcbNativeHandle(lcbDevice *device, uint64_t handle) {
    if (!device) return 0;
    auto r = device->find_any(handle);
    return r ? r->native() : 0;
}

int lcbBlobInfo(lcbDevice *device, lcrtStream stream, uint64_t handle, lcrtBlobInfo *out) {
    LCB_DEVICE_GUARD(device);
    auto r = device->find_any(handle);
    if (!r || !r->blob) return LCRT_ERROR_INVALID_VALUE;
    return lcrtBlobGetInfo(stream, r->blob, out);
}

int lcbTraceClosest(lcbDevice *device, lcrtStream stream, uint64_t accel, const void *rays, uint64_t ray_count, const uint32_t *masks,
                    uint32_t mask, void *hits, uint64_t *counters) {
    LCB_DEVICE_GUARD(device);
    auto a = device->find<lcb::Accel>(accel, lcb::Resource::Tag::ACCEL);
    if (!a || !a->built) return LCRT_ERROR_INVALID_VALUE;
    device->note_stream(static_cast<cudaStream_t>(stream));
    a->check_pending = true;// a traversal can flag the TLAS header (trace_error)
    return lcrtTraceClosest(stream, a->blob, a->instance_buffer, rays, ray_count, masks, mask, hits, counters);
}

int lcbTraceClosestTiled(lcbDevice *device, lcrtStream stream, uint64_t accel, const void *rays, uint64_t ray_count, const uint32_t *masks,
                         uint32_t mask, void *scratch_hits, void *dst, uint64_t tile_rays, uint64_t tile_stride_bytes, uint32_t flags,
                         uint64_t *counters) {
    LCB_DEVICE_GUARD(device);
    auto a = device->find<lcb::Accel>(accel, lcb::Resource::Tag::ACCEL);
    if (!a || !a->built) return LCRT_ERROR_INVALID_VALUE;
    device->note_stream(static_cast<cudaStream_t>(stream));
    a->check_pending = true;
    return lcrtTraceClosestTiled(stream, a->blob, a->instance_buffer, rays, ray_count, masks, mask, scratch_hits, dst, tile_rays, tile_stride_bytes,
                                 flags, counters);
}

// ---- framebuffers shared between the processes of a node (one process per GPU): CUDA IPC over plain cudaMalloc memory
int lcbFramebufferCreate(lcbDevice *device, uint64_t bytes, uint64_t *address, unsigned char handle[64]) {
    LCB_DEVICE_GUARD(device);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (!address || !handle || bytes == 0) return LCRT_ERROR_INVALID_VALUE;
    void *p = nullptr;
    LCB_CUDA(cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        LCB_CUDA(e);
    }
    memcpy(handle, &h, 64);
    *address = reinterpret_cast<uint64_t>(p);
    return LCRT_SUCCESS;
}

int lcbFramebufferOpen(lcbDevice *device, const unsigned char handle[64], uint64_t *address) {
    LCB_DEVICE_GUARD(device);
    if (!address || !handle) return LCRT_ERROR_INVALID_VALUE;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void *p = nullptr;
    LCB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *address = reinterpret_cast<uint64_t>(p);
    return LCRT_SUCCESS;
}

int lcbFramebufferClose(lcbDevice *device, uint64_t address) {
    LCB_DEVICE_GUARD(device);
    LCB_CUDA(cudaIpcCloseMemHandle(reinterpret_cast<void *>(address)));
    return LCRT_SUCCESS;
}

int lcbFramebufferCopyTiles(lcbDevice *device, lcrtStream stream, const void *src, uint64_t dst, uint64_t tile_bytes, uint64_t dst_stride_bytes,
                            uint64_t tile_count) {
    LCB_DEVICE_GUARD(device);
    if (tile_count == 0) return LCRT_SUCCESS;
    if (!src || !dst || tile_bytes == 0 || dst_stride_bytes < tile_bytes) return LCRT_ERROR_INVALID_VALUE;
    LCB_CUDA(cudaMemcpy2DAsync(reinterpret_cast<void *>(dst), dst_stride_bytes, src, tile_bytes, tile_bytes, tile_count, cudaMemcpyDeviceToDevice,
                               static_cast<cudaStream_t>(stream)));
    return LCRT_SUCCESS;
}

int lcbFramebufferDestroy(lcbDevice *device, uint64_t address) {
    LCB_DEVICE_GUARD(device);
    LCB_CUDA(cudaFree(reinterpret_cast<void *>(address)));
    return LCRT_SUCCESS;
}

int lcbTraceAny(lcbDevice *device, lcrtStream stream, uint64_t accel, const void *rays, uint64_t ray_count, const uint32_t *masks,
                uint32_t mask, uint32_t *occluded, uint64_t *counters) {
    LCB_DEVICE_GUARD(device);
    auto a = device->find<lcb::Accel>(accel, lcb::Resource::Tag::ACCEL);
    if (!a || !a->built) return LCRT_ERROR_INVALID_VALUE;
    device->note_stream(static_cast<cudaStream_t>(stream));
    a->check_pending = true;// a traversal can flag the TLAS header (trace_error)
    return lcrtTraceAny(stream, a->blob, a->instance_buffer, rays, ray_count, masks, mask, occluded, counters);
}

int lcbQuery(lcbDevice *device, lcrtStream stream, uint64_t accel, const void *rays, uint64_t ray_count, const uint32_t *masks, uint32_t mask,
             int terminate_on_first, uint32_t tri_program, uint32_t proc_program, void *hits) {
    LCB_DEVICE_GUARD(device);
    auto a = device->find<lcb::Accel>(accel, lcb::Resource::Tag::ACCEL);
    if (!a || !a->built) return LCRT_ERROR_INVALID_VALUE;
    device->note_stream(static_cast<cudaStream_t>(stream));
    a->check_pending = true;// a traversal can flag the TLAS header (trace_error)
    return lcrtQuery(stream, a->blob, a->instance_buffer, rays, ray_count, masks, mask, terminate_on_first, tri_program, proc_program, hits);
}

int lcbTraceClosestHost(lcbDevice *device, uint64_t accel, const void *host_rays, uint64_t ray_count, uint32_t mask, void *host_hits,
                        uint64_t chunk_rays) {
    LCB_DEVICE_GUARD(device);
    auto a = device->find<lcb::Accel>(accel, lcb::Resource::Tag::ACCEL);
    if (!a || !a->built || (!host_rays && ray_count) || (!host_hits && ray_count)) return LCRT_ERROR_INVALID_VALUE;
    if (chunk_rays == 0) chunk_rays = 1u << 20;
    constexpr int SLOTS = lcbDevice::STAGE_SLOTS;
    if (device->stage_capacity < chunk_rays) {
        for (int i = 0; i < SLOTS; i++) {
            if (device->stage_rays[i]) cudaFree(device->stage_rays[i]);
            if (device->stage_hits[i]) cudaFree(device->stage_hits[i]);
            LCB_CUDA(cudaMalloc(&device->stage_rays[i], chunk_rays * sizeof(lcrt::Ray)));
            LCB_CUDA(cudaMalloc(&device->stage_hits[i], chunk_rays * sizeof(lcrt::SurfaceHit)));
            if (!device->io_streams[i]) LCB_CUDA(cudaStreamCreateWithFlags(&device->io_streams[i], cudaStreamNonBlocking));
        }
        device->stage_capacity = chunk_rays;
    }
    // the internal streams start behind the builds this trace depends on (they may still be running on the caller's streams)
    {
        std::scoped_lock lock{a->mutex};
        for (int i = 0; i < SLOTS; i++) {
            if (a->ready) LCB_CUDA(cudaStreamWaitEvent(device->io_streams[i], a->ready, 0));
            for (auto &kv : a->held)
                if (kv.second.prim->ready) LCB_CUDA(cudaStreamWaitEvent(device->io_streams[i], kv.second.prim->ready, 0));
        }
    }
    // SLOTS streams, SLOTS staging slots: chunk k's H2D / trace / D2H are ordered on stream k % SLOTS.  With four
    // chunks in flight the host->device copy engine always has the next chunk queued while earlier chunks are being
    // traversed and read back, so PCIe runs in both directions at once and the traversal hides under the copies
    // (measured on the B200 box: 49 GB/s per direction with both directions busy; tools/pcie_probe.py).
    uint64_t done = 0;
    for (uint64_t k = 0; done < ray_count; k++) {
        const int s = (int)(k % SLOTS);
        const uint64_t n = ray_count - done < chunk_rays ? ray_count - done : chunk_rays;
        cudaStream_t st = device->io_streams[s];
        LCB_CUDA(cudaMemcpyAsync(device->stage_rays[s], static_cast<const char *>(host_rays) + done * sizeof(lcrt::Ray), n * sizeof(lcrt::Ray),
                                 cudaMemcpyHostToDevice, st));
        LCB_CHECK(lcrtTraceClosest(st, a->blob, a->instance_buffer, device->stage_rays[s], n, nullptr, mask, device->stage_hits[s], nullptr));
        LCB_CUDA(cudaMemcpyAsync(static_cast<char *>(host_hits) + done * sizeof(lcrt::SurfaceHit), device->stage_hits[s], n * sizeof(lcrt::SurfaceHit),
                                 cudaMemcpyDeviceToHost, st));
        done += n;
    }
    for (int i = 0; i < SLOTS; i++) LCB_CUDA(cudaStreamSynchronize(device->io_streams[i]));
    return 0;
}

int lcbPrimitiveExport(lcbDevice *device, lcrtStream stream, uint64_t handle, void *dst_device, uint64_t capacity, uint64_t *bytes) {
    LCB_DEVICE_GUARD(device);
    auto r = device->find_any(handle);
    if (!r || r->tag == lcb::Resource::Tag::ACCEL || !r->blob || !bytes) return LCRT_ERROR_INVALID_VALUE;
    std::scoped_lock lock{r->mutex};
    *bytes = r->blob_bytes;
    if (dst_device == nullptr) return 0;
    if (capacity < r->blob_bytes) return LCRT_ERROR_BUFFER_TOO_SMALL;
    LCB_CUDA(cudaMemcpyAsync(dst_device, r->blob, r->blob_bytes, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
    return 0;
}

int lcbPrimitiveImport(lcbDevice *device, lcrtStream stream, uint64_t handle, const void *src_device, uint64_t bytes) {
    LCB_DEVICE_GUARD(device);
    auto r = device->find_any(handle);
    if (!r || r->tag == lcb::Resource::Tag::ACCEL || !src_device || bytes < sizeof(lcrt::BlobHeader)) return LCRT_ERROR_INVALID_VALUE;
    std::scoped_lock lock{r->mutex};
    auto st = static_cast<cudaStream_t>(stream);
    device->note_stream(st);
    // the header says what arrives (import is rare: one small synchronous read-back)
    lcrt::BlobHeader h;
    LCB_CUDA(cudaMemcpyAsync(&h, src_device, sizeof(h), cudaMemcpyDeviceToHost, st));
    LCB_CUDA(cudaStreamSynchronize(st));
    const uint32_t want = r->tag == lcb::Resource::Tag::MESH ? lcrt::KIND_TRIANGLES : lcrt::KIND_AABBS;
    if (h.magic != lcrt::BLOB_MAGIC || !h.built || h.error != 0u || h.kind != want || h.total_bytes > bytes) {
        fprintf(stderr, "[lcb200] lcbPrimitiveImport: not a built blob of this primitive kind (magic %08x, kind %u, built %u, error %u)\n", h.magic, h.kind,
                h.built, h.error);
        return LCRT_ERROR_INVALID_VALUE;
    }
    if (r->blob) LCB_CUDA(cudaFreeAsync(r->blob, st));
    r->blob = nullptr, r->blob_bytes = 0;
    LCB_CUDA(cudaMallocAsync(&r->blob, bytes, st));
    r->blob_bytes = bytes;
    device->blob_epoch.fetch_add(1);
    LCB_CUDA(cudaMemcpyAsync(r->blob, src_device, bytes, cudaMemcpyDeviceToDevice, st));
    auto prim = static_cast<lcb::Primitive *>(r);
    prim->built = true, prim->blob_updatable = h.parent_offset != 0ull;
    r->last_build_kind = 1;
    return r->mark_ready(st);
}

int lcbSynchronize(lcbDevice *device, lcrtStream stream) {
    LCB_DEVICE_GUARD(device);
    LCB_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    // builds never stall the host, so their device-side status is looked at here, at the caller's own synchronisation point
    return device->surface_build_errors();
}

int lcbDownload(lcbDevice *device, lcrtStream stream, uint64_t device_address, void *host, uint64_t bytes) {
    LCB_DEVICE_GUARD(device);
    LCB_CUDA(cudaMemcpyAsync(host, reinterpret_cast<const void *>(device_address), bytes, cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
    LCB_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    return 0;
}

int lcbUpload(lcbDevice *device, lcrtStream stream, uint64_t device_address, const void *host, uint64_t bytes) {
    LCB_DEVICE_GUARD(device);
    LCB_CUDA(cudaMemcpyAsync(reinterpret_cast<void *>(device_address), host, bytes, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
    LCB_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    return 0;
}

// ---- plumbing for the backend plugin (plugin/lc_backend_b200rt.cpp): device memory, streams, copies, host callbacks
// and timeline events, so that the plugin itself contains no CUDA calls (cuda_device.cpp:281-340, cuda_stream.cpp,
// cuda_event.cpp are what it stands in for)
uint64_t lcbBufferCreate(lcbDevice *device, uint64_t bytes) {
    if (!device) return ~0ull;
    cudaSetDevice(device->index);
    void *p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) return ~0ull;
    return reinterpret_cast<uint64_t>(p);
}
void lcbBufferDestroy(lcbDevice *device, uint64_t address) {
    if (!device || address == ~0ull) return;
    cudaSetDevice(device->index);
    cudaFree(reinterpret_cast<void *>(address));
}
lcrtStream lcbStreamCreate(lcbDevice *device) {
    if (!device) return nullptr;
    cudaSetDevice(device->index);
    cudaStream_t s = nullptr;
    if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    return s;
}
void lcbStreamDestroy(lcbDevice *device, lcrtStream stream) {
    if (!device || !stream) return;
    cudaSetDevice(device->index);
    cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
    device->forget_stream(static_cast<cudaStream_t>(stream));
    cudaStreamDestroy(static_cast<cudaStream_t>(stream));
}
int lcbCopyAsync(lcbDevice *device, lcrtStream stream, void *dst, const void *src, uint64_t bytes) {
    LCB_DEVICE_GUARD(device);
    LCB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, static_cast<cudaStream_t>(stream)));
    return 0;
}
int lcbHostCallback(lcbDevice *device, lcrtStream stream, void (*fn)(void *), void *user) {
    LCB_DEVICE_GUARD(device);
    LCB_CUDA(cudaLaunchHostFunc(static_cast<cudaStream_t>(stream), fn, user));
    return 0;
}
// timeline event: signal(value) records a CUDA event tagged with the fence value, wait / synchronise use the first
// recorded event whose value is >= the one asked for
struct lcbEvent {
    std::mutex mutex;
    std::vector<std::pair<uint64_t, cudaEvent_t>> signals;
};
lcbEvent *lcbEventCreate(lcbDevice *device) { return device ? new lcbEvent{} : nullptr; }
void lcbEventDestroy(lcbDevice *device, lcbEvent *event) {
    if (!device || !event) return;
    cudaSetDevice(device->index);
    for (auto &s : event->signals) cudaEventDestroy(s.second);
    delete event;
}
int lcbEventSignal(lcbDevice *device, lcbEvent *event, lcrtStream stream, uint64_t value) {
    LCB_DEVICE_GUARD(device);
    if (!event) return LCRT_ERROR_INVALID_VALUE;
    cudaEvent_t e;
    LCB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    LCB_CUDA(cudaEventRecord(e, static_cast<cudaStream_t>(stream)));
    std::scoped_lock lock{event->mutex};
    // completed signals below the new value are no longer needed
    for (auto it = event->signals.begin(); it != event->signals.end();) {
        if (it->first < value && cudaEventQuery(it->second) == cudaSuccess) {
            cudaEventDestroy(it->second);
            it = event->signals.erase(it);
        } else {
            ++it;
        }
    }
    event->signals.emplace_back(value, e);
    return 0;
}
static cudaEvent_t lcb_event_for(lcbEvent *event, uint64_t value, bool *known) {
    std::scoped_lock lock{event->mutex};
    cudaEvent_t best = nullptr;
    uint64_t bv = ~0ull;
    *known = false;
    for (auto &s : event->signals)
        if (s.first >= value && s.first <= bv) best = s.second, bv = s.first, *known = true;
    return best;
}
int lcbEventWait(lcbDevice *device, lcbEvent *event, lcrtStream stream, uint64_t value) {
    LCB_DEVICE_GUARD(device);
    if (!event) return LCRT_ERROR_INVALID_VALUE;
    bool known;
    cudaEvent_t e = lcb_event_for(event, value, &known);
    if (!known) return LCRT_ERROR_INVALID_VALUE;// waiting on a value nobody has signalled yet (the reference asserts this too)
    LCB_CUDA(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), e, 0));
    return 0;
}
int lcbEventSynchronize(lcbDevice *device, lcbEvent *event, uint64_t value) {
    LCB_DEVICE_GUARD(device);
    if (!event) return LCRT_ERROR_INVALID_VALUE;
    bool known;
    cudaEvent_t e = lcb_event_for(event, value, &known);
    if (!known) return value == 0 ? 0 : LCRT_ERROR_INVALID_VALUE;
    LCB_CUDA(cudaEventSynchronize(e));
    return 0;
}
int lcbEventIsCompleted(lcbDevice *device, lcbEvent *event, uint64_t value) {
    if (!device || !event) return 0;
    cudaSetDevice(device->index);
    bool known;
    cudaEvent_t e = lcb_event_for(event, value, &known);
    if (!known) return value == 0 ? 1 : 0;
    return cudaEventQuery(e) == cudaSuccess ? 1 : 0;
}

}// extern "C"
