"""CPU checks of the Java/JNI side of the drop-in boundary (the image has no JDK, so javac cannot run):

* jni/gsql_jni.c compiles with -Wall -Werror against jni/stub/jni.h (the JNI specification's names and signatures)
  and LINKS against libgsql_gpu.so with --no-undefined: every gsql_* call it makes exists with that arity;
* every `native` method of GpuNative.java has its Java_..._GpuNative_<name> definition in the shim, with the same
  number of parameters, and the shim defines nothing else;
* every Gpu* class the Java sources reference exists in java/, every GpuNative.<method>(...) call site names a declared
  native, and the operators that override AbstractExecutor's package-private template methods live in its package."""
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JAVA = os.path.join(ROOT, "java")
PKG = "com/alibaba/polardbx/executor"


def _java_files():
    out = {}
    for d, _, files in os.walk(JAVA):
        for f in files:
            if f.endswith(".java"):
                out[os.path.join(d, f)] = open(os.path.join(d, f)).read()
    return out


def _strip_comments(src):
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return re.sub(r"//[^\n]*", "", src)


def _natives():
    src = _strip_comments(open(os.path.join(JAVA, PKG, "operator/gpu/GpuNative.java")).read())
    out = {}
    for m in re.finditer(r"public static native\s+[\w\[\]]+\s+(\w+)\s*\(([^)]*)\)\s*;", src, flags=re.S):
        params = [p for p in m.group(2).split(",") if p.strip()]
        out[m.group(1)] = len(params)
    return out


def test_jni_shim_compiles_and_links_against_the_abi():
    import __graft_entry__ as g
    g.build()
    so_dir = os.path.join(ROOT, "galaxysql_b200", "_build")
    with tempfile.TemporaryDirectory() as d:
        cmd = ["gcc", "-shared", "-fPIC", "-Wall", "-Werror", "-I", os.path.join(ROOT, "jni", "stub"), "-I", os.path.join(ROOT, "include"),
               os.path.join(ROOT, "jni", "gsql_jni.c"), "-L", so_dir, "-lgsql_gpu", "-Wl,--no-undefined", "-o", os.path.join(d, "libgsql_jni.so")]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        syms = subprocess.check_output(["nm", "-D", "--defined-only", os.path.join(d, "libgsql_jni.so")], text=True)
    exported = set(re.findall(r"Java_com_alibaba_polardbx_executor_operator_gpu_GpuNative_(\w+)", syms))
    assert exported == set(_natives()), (sorted(set(_natives()) - exported), sorted(exported - set(_natives())))


def test_every_native_has_a_definition_with_matching_arity():
    natives = _natives()
    assert len(natives) >= 30
    src = _strip_comments(open(os.path.join(ROOT, "jni", "gsql_jni.c")).read())
    defs = {}
    for m in re.finditer(r"NATIVE\(\s*\w+\s*,\s*(\w+)\s*\)\s*\(([^)]*)\)", src, flags=re.S):
        defs[m.group(1)] = len([p for p in m.group(2).split(",") if p.strip()]) - 2   # JNIEnv *, jclass
    assert set(defs) == set(natives)
    for name, n in natives.items():
        assert defs[name] == n, f"{name}: Java declares {n} parameters, the shim takes {defs[name]}"


def test_java_sources_are_closed_over_their_gpu_classes():
    files = _java_files()
    classes = {os.path.basename(p)[:-5] for p in files}
    natives = _natives()
    for path, raw in files.items():
        src = _strip_comments(raw)
        for ref in set(re.findall(r"\b(Gpu[A-Z]\w*)\b", src)):
            base = ref.split(".")[0]
            assert base in classes or base == "GpuJoinShared", f"{path} references {ref}, which has no source file"
        for m in re.finditer(r"GpuNative\.(\w+)\s*\(", src):
            name = m.group(1)
            assert name in natives or name.startswith("T_"), f"{path} calls GpuNative.{name}, which is not declared"
        # arity of the call sites (top-level commas of the argument list)
        for m in re.finditer(r"GpuNative\.(\w+)\s*\(", src):
            name, i, depth, args, cur = m.group(1), m.end(), 1, 0, ""
            while depth and i < len(src):
                ch = src[i]
                if ch in "([{":
                    depth += 1
                elif ch in ")]}":
                    depth -= 1
                elif ch == "," and depth == 1:
                    args += 1
                if depth:
                    cur += ch
                i += 1
            nargs = 0 if not cur.strip() else args + 1
            assert nargs == natives[name], f"{path}: GpuNative.{name} called with {nargs} arguments, declared with {natives[name]}"
        pkg = re.search(r"^package\s+([\w.]+);", src, flags=re.M).group(1)
        assert path.endswith(os.path.join(*pkg.split(".")) + os.sep + os.path.basename(path)), f"{path}: package {pkg} does not match its directory"
        if re.search(r"\bvoid\s+doOpen\s*\(|\bChunk\s+doNextChunk\s*\(|\bvoid\s+doClose\s*\(", src):
            # package-private template methods of AbstractExecutor (AbstractExecutor.java:87-91): same package or it cannot compile
            assert pkg == "com.alibaba.polardbx.executor.operator", f"{path} overrides package-private hooks from package {pkg}"
            assert "public void doOpen" not in src and "protected void doOpen" not in src


def test_reference_types_the_java_side_imports_exist_where_cited():
    """The imports name reference classes by package; spot-check the ones the design hinges on against INTEGRATION.md's
    citations (paths only — /root/reference is not readable from the tests at run time on the GPU box, so the list is static)."""
    files = _java_files()
    need = {"com.alibaba.polardbx.executor.operator.AbstractExecutor": "operator/GpuHashAggExec.java",
            "com.alibaba.polardbx.executor.mpp.operator.factory.ExecutorFactory": "mpp/operator/factory/GpuHashAggExecutorFactory.java",
            "com.alibaba.polardbx.executor.mpp.operator.LocalExchanger": "mpp/operator/GpuPartitioningExchanger.java"}
    for cls, rel in need.items():
        src = files[os.path.join(JAVA, PKG, rel)]
        simple = cls.split(".")[-1]
        pkg = ".".join(cls.split(".")[:-1])
        assert re.search(r"extends\s+" + simple + r"\b", src), f"{rel} must extend {simple}"
        assert f"package {pkg};" in src or f"import {cls};" in src, f"{rel}: {simple} must be visible (same package or imported)"


def _header_enum_values():
    """name -> value of every `NAME = <int>` enumerator / `#define NAME <int>` in include/gsql_gpu.h."""
    src = _strip_comments(open(os.path.join(ROOT, "include", "gsql_gpu.h")).read())
    out = {m.group(1): int(m.group(2)) for m in re.finditer(r"\b(GSQL_[A-Z0-9_]+)\s*=\s*(-?\d+)", src)}
    out.update({m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+(GSQL_[A-Z0-9_]+)\s+(\d+)\b", src)})
    return out


def test_java_constants_mirror_the_header():
    """The Java side restates a few enumerations of include/gsql_gpu.h as int constants (no generated binding in an image
    without a JDK): type codes, aggregate kinds, expression opcodes, the instruction limit.  They must stay equal."""
    h = _header_enum_values()
    files = _java_files()

    def consts(rel):
        src = _strip_comments(files[os.path.join(JAVA, PKG, rel)])
        out = {}
        for decl in re.findall(r"public static final int\s+([^;]+);", src, flags=re.S):
            for name, val in re.findall(r"(\w+)\s*=\s*(-?\d+)", decl):
                out[name] = int(val)
        return out

    native = consts("operator/gpu/GpuNative.java")
    for name in ("T_INT32", "T_INT64", "T_FP64", "T_DEC128"):
        assert native[name] == h["GSQL_" + name], name
    agg = consts("operator/gpu/GpuAggSpec.java")
    for name in ("COUNT_STAR", "COUNT", "SUM", "AVG", "MIN", "MAX", "SUM0"):
        assert agg[name] == h["GSQL_AGG_" + name], name
    expr = consts("operator/gpu/GpuExpression.java")
    ops = {k: v for k, v in expr.items() if k.startswith("OP_")}
    assert len(ops) >= 19
    for name, val in ops.items():
        assert val == h["GSQL_" + name], name
    assert expr["MAX_INSTRUCTIONS"] == h["GSQL_MAX_EXPR_INS"]
    # every opcode of the header that a program may contain has its Java constant
    assert {k[len("GSQL_"):] for k in h if k.startswith("GSQL_OP_")} == set(ops)


def test_pages_serde_frame_layout_matches_the_codec():
    """GpuPagesSerde splits / builds the 13-byte page frame by hand; the offsets must be the oracle codec's
    (oracle/serde.py restates PagesSerdeUtil.writeSerializedChunk): int32 positionCount | int8 marker | int32 uncompressedSize |
    int32 sizeInBytes, little-endian."""
    import numpy as np
    from oracle import serde
    cols = [(np.arange(5, dtype=np.int64), None), (np.arange(5, dtype=np.float64), np.array([0, 1, 0, 0, 1], bool))]
    blob = serde.serialize(cols, [1, 2], 5)
    pos, marker, unc, size = int.from_bytes(blob[0:4], "little"), blob[4], int.from_bytes(blob[5:9], "little"), int.from_bytes(blob[9:13], "little")
    assert (pos, marker) == (5, 0) and unc == size == len(blob) - 13
    src = _strip_comments(open(os.path.join(JAVA, PKG, "mpp/execution/buffer/GpuPagesSerde.java")).read())
    assert "FRAME_BYTES = 4 + 1 + 4 + 4" in src
    for needle in ("all.getInt(at)", "all.getInt(at + 5)", "all.getInt(at + 9)", "f.setInt(0,", "f.setByte(4, 0)", "f.setInt(5,", "f.setInt(9,"):
        assert needle in src, needle
