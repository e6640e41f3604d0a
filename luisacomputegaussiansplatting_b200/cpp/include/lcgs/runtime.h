// lcgs/runtime.h -- the sliver of a compute runtime the lcgs classes need, over CUDA directly.
//
// The reference builds on LuisaCompute's Context/Device/Stream/Buffer<T>/BufferView<T>/CommandList
// (used by app/main.cpp:43,162-163,180-186,216-223,232-254,313-315).  This facade keeps those names
// and call shapes so that code written against the reference's lcgs API reads the same, but there
// is one backend (CUDA on sm_100a), no JIT and no DSL: "commands" execute on the stream as they are
// appended.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "lcgs_b200.h"

namespace lcgs
{

using uint  = uint32_t;
using ulong = uint64_t;

struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
struct uint2 { uint x, y; };

inline float3 make_float3(float x, float y, float z) { return { x, y, z }; }
inline float3 make_float3(float v) { return { v, v, v }; }
inline uint2  make_uint2(uint x, uint y) { return { x, y }; }

[[noreturn]] inline void fatal(const std::string& what)
{
    // the reference's LUISA_ERROR aborts; so do we (every lcgs entry point is noexcept upstream)
    std::fprintf(stderr, "[lcgs] fatal: %s\n", what.c_str());
    std::abort();
}

inline void cuda_check(cudaError_t e, const char* what)
{
    if (e != cudaSuccess) fatal(std::string(what) + ": " + cudaGetErrorString(e));
}

// A unit of work that is enqueued on the stream its command list is committed to: the stand-in for
// luisa::compute::Command.  As in LuisaCompute, `cmdlist << command` only RECORDS the command;
// `stream << cmdlist.commit()` enqueues the recorded commands in order and `<< synchronize()` waits.
using Command = std::function<void(cudaStream_t)>;

template <typename T>
struct BufferView {
    T*     ptr   = nullptr;
    size_t count = 0;
    size_t       size() const noexcept { return count; }
    T*           data() const noexcept { return ptr; }
    BufferView   subview(size_t offset, size_t n) const noexcept { return { ptr + offset, n }; }
    explicit     operator bool() const noexcept { return ptr != nullptr; }
    // `cmdlist << view.copy_to(host)` / `<< view.copy_from(host)` (app/main.cpp:216-222,313-315, impl.cpp:106)
    Command copy_to(void* host) const
    {
        return [p = ptr, n = count, host](cudaStream_t s) {
            cuda_check(cudaMemcpyAsync(host, p, n * sizeof(T), cudaMemcpyDeviceToHost, s), "copy_to");
        };
    }
    Command copy_from(const void* host) const
    {
        return [p = ptr, n = count, host](cudaStream_t s) {
            cuda_check(cudaMemcpyAsync(p, host, n * sizeof(T), cudaMemcpyHostToDevice, s), "copy_from");
        };
    }
};

class Stream;

// Owning device allocation.  copy_from / copy_to return "commands" (closures over the stream) to
// mirror `cmd_list << buf.copy_from(host)`.
template <typename T>
class Buffer
{
public:
    Buffer() = default;
    explicit Buffer(size_t n) : m_n(n)
    {
        cuda_check(cudaMalloc(&m_ptr, (n ? n : 1) * sizeof(T)), "cudaMalloc");
        // zero-initialised: defines the contents the reference leaves stale
        cuda_check(cudaMemset(m_ptr, 0, (n ? n : 1) * sizeof(T)), "cudaMemset");
    }
    Buffer(const Buffer&)            = delete;
    Buffer& operator=(const Buffer&) = delete;
    Buffer(Buffer&& o) noexcept : m_ptr(o.m_ptr), m_n(o.m_n) { o.m_ptr = nullptr; o.m_n = 0; }
    Buffer& operator=(Buffer&& o) noexcept
    {
        if (this != &o) { release(); m_ptr = o.m_ptr; m_n = o.m_n; o.m_ptr = nullptr; o.m_n = 0; }
        return *this;
    }
    ~Buffer() { release(); }

    size_t        size() const noexcept { return m_n; }
    BufferView<T> view() const noexcept { return { m_ptr, m_n }; }
    BufferView<T> view(size_t offset, size_t n) const noexcept { return { m_ptr + offset, n }; }
    BufferView<T> subview(size_t offset, size_t n) const noexcept { return view(offset, n); }
    operator BufferView<T>() const noexcept { return view(); }
    Command       copy_to(void* host) const { return view().copy_to(host); }
    Command       copy_from(const void* host) const { return view().copy_from(host); }

private:
    void release() noexcept
    {
        if (m_ptr) cudaFree(m_ptr);
        m_ptr = nullptr;
    }
    T*     m_ptr = nullptr;
    size_t m_n   = 0;
};

struct CommitToken {         // what CommandList::commit() hands to `stream << ...`: the recorded commands
    std::vector<Command> commands;
};
struct SynchronizeToken {};  // luisa::compute::synchronize()
inline SynchronizeToken synchronize() noexcept { return {}; }

class Stream
{
public:
    Stream() { cuda_check(cudaStreamCreateWithFlags(&m_s, cudaStreamNonBlocking), "cudaStreamCreate"); }
    Stream(const Stream&)            = delete;
    Stream& operator=(const Stream&) = delete;
    ~Stream() { if (m_s) cudaStreamDestroy(m_s); }
    cudaStream_t     handle() const noexcept { return m_s; }
    lcgs_b200_stream abi() const noexcept { return reinterpret_cast<lcgs_b200_stream>(m_s); }
    void             synchronize() const { cuda_check(cudaStreamSynchronize(m_s), "cudaStreamSynchronize"); }
    // `stream << cmdlist.commit() << synchronize()` (impl.cpp:100,107,131,144,146)
    Stream& operator<<(CommitToken&& t)
    {
        for (auto& c : t.commands) c(m_s);
        return *this;
    }
    Stream& operator<<(SynchronizeToken) { synchronize(); return *this; }
    Stream& operator<<(const Command& c) { c(m_s); return *this; }
    template <typename T>
    void upload(BufferView<T> dst, const T* host) const
    {
        cuda_check(cudaMemcpyAsync(dst.ptr, host, dst.count * sizeof(T), cudaMemcpyHostToDevice, m_s), "H2D");
    }
    template <typename T>
    void download(BufferView<T> src, T* host) const
    {
        cuda_check(cudaMemcpyAsync(host, src.ptr, src.count * sizeof(T), cudaMemcpyDeviceToHost, m_s), "D2H");
    }

private:
    cudaStream_t m_s = nullptr;
};

// luisa::compute::CommandList: records commands until commit() (impl.cpp:87-177 appends kernels, fills, the
// scan, the sort and a read-back to one list and commits it five times).
class CommandList
{
public:
    CommandList() = default;
    CommandList& operator<<(Command c)
    {
        m_commands.push_back(std::move(c));
        return *this;
    }
    CommitToken commit() noexcept
    {
        CommitToken t{ std::move(m_commands) };
        m_commands.clear();
        return t;
    }
    bool empty() const noexcept { return m_commands.empty(); }

private:
    std::vector<Command> m_commands;
};

class Device
{
public:
    explicit Device(int index = 0) : m_index(index)
    {
        const int rc = lcgs_b200_ctx_create(index, &m_ctx);
        if (rc != LCGS_B200_OK) fatal(std::string("lcgs_b200_ctx_create: ") + lcgs_b200_status_string(rc));
        cuda_check(cudaSetDevice(index), "cudaSetDevice");
    }
    Device(const Device&)            = delete;
    Device& operator=(const Device&) = delete;
    ~Device() { lcgs_b200_ctx_destroy(m_ctx); }

    template <typename T>
    Buffer<T> create_buffer(size_t n) const { return Buffer<T>(n); }
    lcgs_b200_ctx* ctx() const noexcept { return m_ctx; }
    int            index() const noexcept { return m_index; }
    void           check(int rc, const char* what) const
    {
        if (rc != LCGS_B200_OK)
            fatal(std::string(what) + ": " + lcgs_b200_status_string(rc) + " (" + lcgs_b200_last_error(m_ctx) + ")");
    }

private:
    lcgs_b200_ctx* m_ctx   = nullptr;
    int            m_index = 0;
};

// Base of the lcgs modules (reference: LuisaModule / GSModule with m_blocks = {16,16}, module.h:17)
class GSModule
{
public:
    uint2 m_blocks = { 16u, 16u };

protected:
    Device* m_device = nullptr;
};

}  // namespace lcgs
