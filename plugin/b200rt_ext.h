// b200rt_ext.h — what a LuisaCompute program includes to trace rays through the `b200rt` backend plugin without a DSL
// kernel: a CustomCommand (the reference's hook for backend-specific work, include/luisa/runtime/rhi/command.h:682-690)
// that traces a buffer of rays against an Accel.  It stands in for a ShaderDispatchCommand of a kernel that only calls
// accel.intersect / intersect_any / query_all / query_any (src/tests/test_rtx.cpp:52-99); record layouts are the
// reference's (Ray 32 B, SurfaceHit / CommittedHit 24 B, include/luisa/runtime/rtx/ray.h:10-16, hit.h:7-43).
#pragma once
#include <luisa/runtime/rhi/command.h>

namespace luisa::compute::b200rt {

constexpr uint64_t trace_command_uuid = 0xb2005254'54524143ull;// "B200RT" "TRAC"

class TraceCommand final : public CustomCommand {
public:
    enum struct Mode : uint32_t {
        CLOSEST = 0,  // accel.intersect      -> SurfaceHit[n]
        ANY = 1,      // accel.intersect_any  -> uint[n] (1 = occluded)
        QUERY_ALL = 2,// accel.query_all with the built-in candidate programs -> CommittedHit[n]
        QUERY_ANY = 3 // accel.query_any
    };

private:
    uint64_t _accel, _rays, _hits;
    size_t _ray_offset, _hit_offset, _count;
    uint32_t _mask;
    Mode _mode;
    uint32_t _triangle_program{0u}, _procedural_program{0u};

public:
    // handles as returned by Accel::handle() / Buffer::handle(); offsets in bytes
    TraceCommand(Mode mode, uint64_t accel, uint64_t ray_buffer, size_t ray_offset_bytes, uint64_t hit_buffer,
                 size_t hit_offset_bytes, size_t ray_count, uint32_t visibility_mask = 0xffu) noexcept
        : _accel{accel}, _rays{ray_buffer}, _hits{hit_buffer}, _ray_offset{ray_offset_bytes}, _hit_offset{hit_offset_bytes},
          _count{ray_count}, _mask{visibility_mask}, _mode{mode} {}
    // built-in candidate programs of the query modes (lcb200_rt.h: LCRT_TRI_PROGRAM_* / LCRT_PROC_PROGRAM_*)
    void set_programs(uint32_t triangle_program, uint32_t procedural_program) noexcept {
        _triangle_program = triangle_program, _procedural_program = procedural_program;
    }
    [[nodiscard]] uint64_t uuid() const noexcept override { return trace_command_uuid; }
    [[nodiscard]] StreamTag stream_tag() const noexcept override { return StreamTag::COMPUTE; }
    [[nodiscard]] auto mode() const noexcept { return _mode; }
    [[nodiscard]] auto accel() const noexcept { return _accel; }
    [[nodiscard]] auto ray_buffer() const noexcept { return _rays; }
    [[nodiscard]] auto hit_buffer() const noexcept { return _hits; }
    [[nodiscard]] auto ray_offset() const noexcept { return _ray_offset; }
    [[nodiscard]] auto hit_offset() const noexcept { return _hit_offset; }
    [[nodiscard]] auto ray_count() const noexcept { return _count; }
    [[nodiscard]] auto visibility_mask() const noexcept { return _mask; }
    [[nodiscard]] auto triangle_program() const noexcept { return _triangle_program; }
    [[nodiscard]] auto procedural_program() const noexcept { return _procedural_program; }
};

}// namespace luisa::compute::b200rt
