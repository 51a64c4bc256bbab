"""Build step of lc-backend-b200: a build-time copy of the reference's CUDA code generator with this repository's RayQuery
lowering spliced in.

    python make_codegen.py <reference>/src/backends/cuda/cuda_codegen_ast.cpp  b200_ray_query_lowering.inl  <out.cpp>

`class CUDACodegenAST::RayQueryLowering { ... };` (the first class definition of that name in the file) becomes

    class CUDACodegenAST::RayQueryLowering {
    public:
        class Outlined { <the reference's members, constructor renamed> };
    #include "b200_ray_query_lowering.inl"
    };

Nothing else of the file changes, nothing of it is kept in this repository: the output lives in the build directory."""
import re
import sys


def main():
    src, inl, out = sys.argv[1:4]
    text = open(src).read()
    head = "class CUDACodegenAST::RayQueryLowering {"
    a = text.index(head)
    m = re.compile(r"^};\s*$", re.M).search(text, a)      # the class is closed by the first `};` in column 0 after it
    assert m, "end of RayQueryLowering not found"
    body = text[a + len(head):m.start()]
    assert body.count("explicit RayQueryLowering(") == 1, "constructor of the reference's RayQueryLowering not found"
    body = body.replace("explicit RayQueryLowering(", "explicit Outlined(")
    assert "lc_ray_query_trace(" in body and "_create_outline_definitions" in body, "this is not the lowering this script knows"
    new = (head + "\npublic:\n    // the reference's lowering (outlined handlers + capture struct), unmodified\n    class Outlined {" + body + "    };\n\n"
           + '#include "' + inl + '"\n')
    text = text[:a] + new + text[m.start():]
    open(out, "w").write(text)
    print(f"{out}: RayQueryLowering spliced ({len(body.splitlines())} reference lines kept as Outlined)")


if __name__ == "__main__":
    main()
