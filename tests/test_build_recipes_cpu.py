"""The three statements of "what the CUDA library is made of" must agree: the Python build recipe
(trueno_b200/build.py), the Rust FFI crate's build.rs (integration/rust/trueno-cuda-sys/build.rs — it cannot be
compiled in this image, so it is pinned by reading) and the directory itself; and the three statements of the C ABI
must agree too: include/trueno_cuda.h, the FFI crate's `extern "C"` block, and `nm -D` of the built library."""
import glob
import importlib.util
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "trueno_b200", "csrc")
SYS = os.path.join(ROOT, "integration", "rust", "trueno-cuda-sys")


def _build_py():
    spec = importlib.util.spec_from_file_location("_trn_build_t", os.path.join(ROOT, "trueno_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _rust_list(name: str) -> list[str]:
    text = open(os.path.join(SYS, "build.rs")).read()
    body = re.search(r"const %s: &\[&str\] = &\[(.*?)\];" % name, text, re.S).group(1)
    return re.findall(r'"([^"]+)"', body)


def test_source_lists_agree_with_the_directory():
    on_disk = sorted(os.path.basename(p) for p in glob.glob(os.path.join(CSRC, "*.cu")))
    assert sorted(_build_py().SOURCES) == on_disk
    assert _rust_list("SOURCES") == list(_build_py().SOURCES)


def test_build_rs_tracks_every_shared_header_and_links_strictly():
    headers = sorted(os.path.basename(p) for p in glob.glob(os.path.join(CSRC, "*.cuh")))
    assert sorted(_rust_list("HEADERS")) == headers
    text = open(os.path.join(SYS, "build.rs")).read()
    assert "--no-undefined" in text and "trueno_cuda.h" in text
    assert "arch=compute_100a,code=sm_100a" in text and "sm_90" not in text
    # build.py depends on the same headers
    src = open(os.path.join(ROOT, "trueno_b200", "build.py")).read()
    for h in headers:
        assert h in src, f"{h} is not a dependency in trueno_b200/build.py"


def test_header_ffi_crate_and_library_export_the_same_symbols():
    header = open(os.path.join(ROOT, "include", "trueno_cuda.h")).read()
    declared = set(re.findall(r"TRN_API\s+[\w\s\*]+?\b(trn_\w+)\s*\(", header))
    rust = open(os.path.join(SYS, "src", "lib.rs")).read()
    bound = set(re.findall(r"\bpub fn (trn_\w+)\s*\(", rust))
    assert declared == bound, sorted(declared ^ bound)
    lib = os.path.join(ROOT, "trueno_b200", "libtrueno_cuda.so")
    out = subprocess.run(["nm", "-D", "--defined-only", lib], check=True, stdout=subprocess.PIPE, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if ln.split() and ln.split()[-1].startswith("trn_")}
    assert exported == declared, sorted(exported ^ declared)
