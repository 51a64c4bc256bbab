// plugin/hybrid/b200_ray_query_lowering.inl — this repository's lowering of RayQueryStmt for the CUDA code generator.
//
// Spliced by plugin/hybrid/make_codegen.py into a build-time copy of the reference's src/backends/cuda/cuda_codegen_ast.cpp as
// the body of `class CUDACodegenAST::RayQueryLowering` (declared in cuda_codegen_ast.h:20, owned by the generator as
// `_ray_query_lowering`, called from emit() :1310-1312 / :1390 and visit(const RayQueryStmt *) :1282-1284).  The reference's
// own class (cuda_codegen_ast.cpp:45-618) is kept next to it as the nested class `Outlined`, unmodified, and serves kernels
// when LUISA_B200_RAY_QUERY=outline is set.
//
// The reference outlines the two candidate handlers of every RayQuery into functions that OptiX calls as intersection /
// any-hit programs, and moves every variable they touch through a per-site struct behind a `void *` (LCRayQueryCtx<i>,
// cuda_codegen_ast.cpp:255-543 and 562-617): copy in before the trace, copy in / out around every handler call, copy out after.
// Without OptiX nothing needs the handlers to be functions.  This lowering emits the resumable loop the HLSL generator emits
// (src/backends/common/hlsl/hlsl_codegen.cpp:555-567):
//
//     { // ray query #i, lowered inline
//       LCRayQueryInline lc_rq;
//       lc_ray_query_inline_begin(lc_rq, q);
//       while (lc_ray_query_inline_proceed(lc_rq)) {
//         LCIntersectionResult result{};
//         const LCRTCandidate &lc_rt_candidate = lc_rq.candidate;
//         if (lc_rq.triangle_candidate) { <on_triangle_candidate> } else { <on_procedural_candidate> }
//         lc_ray_query_inline_resolve(lc_rq, result);
//       }
//       lc_ray_query_inline_end(lc_rq, q);
//     }
//
// The handler bodies are emitted by the generator's ordinary statement visitor in the scope of the enclosing function, so they
// use its variables directly: no capture analysis, no context struct, no copies.  The LC_RAY_QUERY_* builtins inside them
// (cuda_codegen_ast.cpp:998-1004) expand, as in the outlined form, to accesses of `result` and `lc_rt_candidate`
// (luisacompute_b200/csrc/lcrt_device_library.cuh, which also defines the four lc_ray_query_inline_* functions).
private:
    CUDACodegenAST *_codegen;
    Outlined _outlined;
    bool _inline;
    uint _next_site{0u};

    [[nodiscard]] static bool _choose_inline() noexcept {
        // default: inline.  LUISA_B200_RAY_QUERY=outline selects the reference's outlined handlers (kept for comparison).
        if (auto e = getenv("LUISA_B200_RAY_QUERY")) { return luisa::string_view{e} != "outline"; }
        return true;
    }

    [[nodiscard]] static uint _count_sites(Function f) noexcept {
        auto n = 0u;
        traverse_expressions<true>(
            f.body(),
            [&n](auto expr) noexcept {
                if (expr->tag() == Expression::Tag::CALL) {
                    if (auto call = static_cast<const CallExpr *>(expr); !call->is_builtin()) { n += _count_sites(call->custom()); }
                }
            },
            [&n](auto s) noexcept {
                if (s->tag() == Statement::Tag::RAY_QUERY) { n++; }
            },
            [](auto) noexcept {});
        return n;
    }

public:
    explicit RayQueryLowering(CUDACodegenAST *codegen) noexcept
        : _codegen{codegen}, _outlined{codegen}, _inline{_choose_inline()} {}

    void preprocess(Function f) noexcept {
        if (!_inline) { return _outlined.preprocess(f); }
        // no outlined handlers for the device library to declare; the inline API instead.  An upper bound of the number of
        // sites decides whether lc_ray_query_inline_proceed is inlined (one site) or shared out of line (several).
        _codegen->_scratch << "#define LUISA_RAY_QUERY_IMPL_COUNT 0\n"
                           << "#define LCRT_RAY_QUERY_INLINE_LOWERING 1\n"
                           << "#define LCRT_RAY_QUERY_SITE_COUNT " << std::max(_count_sites(f), 1u) << "\n";
    }

    void outline(Function f) noexcept {
        if (!_inline) { _outlined.outline(f); }
    }

    void lower(const RayQueryStmt *stmt) noexcept {
        if (!_inline) { return _outlined.lower(stmt); }
        auto &s = _codegen->_scratch;
        auto site = _next_site++;
        auto line = [this]() noexcept -> decltype(auto) {
            _codegen->_emit_indent();
            return (_codegen->_scratch);
        };
        s << "\n";
        line() << "{ // ray query #" << site << ", lowered inline\n";
        _codegen->_indent++;
        line() << "LCRayQueryInline lc_rq;\n";
        line() << "lc_ray_query_inline_begin(lc_rq, ";
        stmt->query()->accept(*_codegen);
        s << ");\n";
        line() << "while (lc_ray_query_inline_proceed(lc_rq)) {\n";
        _codegen->_indent++;
        line() << "LCIntersectionResult result{};\n";
        line() << "[[maybe_unused]] const LCRTCandidate &lc_rt_candidate = lc_rq.candidate;\n";
        line() << "if (lc_rq.triangle_candidate) ";
        stmt->on_triangle_candidate()->accept(*_codegen);
        s << " else ";
        stmt->on_procedural_candidate()->accept(*_codegen);
        s << "\n";
        line() << "lc_ray_query_inline_resolve(lc_rq, result);\n";
        _codegen->_indent--;
        line() << "}\n";
        line() << "lc_ray_query_inline_end(lc_rq, ";
        stmt->query()->accept(*_codegen);
        s << ");\n";
        _codegen->_indent--;
        line() << "} // ray query #" << site << "\n";
    }
