"""CPU: the C-ABI library builds, loads, and exports every symbol include/aad.h declares.
No compute calls (no GPU here); host-only entry points are exercised."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "aad.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(aad_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported(built_lib):
    from audioanalysisdetector_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(built_lib, n), f"{n} declared in aad.h but not exported"
        assert n in _lib.SYMBOLS, f"{n} has no ctypes prototype in _lib.SYMBOLS"
    assert sorted(_lib.SYMBOLS) == names


def test_version_and_strerror(built_lib):
    assert built_lib.aad_version() == 200
    assert built_lib.aad_strerror(0) == b"ok"
    assert b"workspace" in built_lib.aad_strerror(-4)


def test_params_default_mirrors_reference_defaults(built_lib):
    from audioanalysisdetector_b200 import _lib as L
    p = L.AadParams()
    assert built_lib.aad_params_default(C.byref(p), L.KIND_LOGMEL, 16000) == 0
    assert p.struct_size == C.sizeof(L.AadParams)
    assert (p.n_fft, p.hop_length, p.n_filt, p.center, p.ref_type, p.n_ceps) == (2048, 512, 64, 1, L.REF_UTT_MAX, 0)
    assert built_lib.aad_params_default(C.byref(p), L.KIND_MFCC, 16000) == 0
    assert (p.n_filt, p.n_ceps, p.ref_type, p.top_db) == (128, 13, L.REF_ONE, 80.0)
    assert built_lib.aad_params_default(C.byref(p), L.KIND_LFCC, 16000) == 0
    assert (p.n_fft, p.win_length, p.hop_length, p.n_filt, p.n_ceps) == (512, 400, 160, 24, 13)
    assert p.quantize_i16 == 1 and abs(p.pre_emph - 0.97) < 1e-7 and p.layout == L.LAYOUT_TC
    assert built_lib.aad_params_default(C.byref(p), 99, 16000) == -1
    assert built_lib.aad_params_default(None, 0, 16000) == -1


def test_python_params_match_c_defaults(built_lib):
    from audioanalysisdetector_b200 import _lib as L
    from audioanalysisdetector_b200.frontend import FrontendParams
    for kind, fp in [(L.KIND_LOGMEL, FrontendParams.logmel(16000)), (L.KIND_MFCC, FrontendParams.mfcc(16000)),
                     (L.KIND_LFCC, FrontendParams.lfcc(16000))]:
        c = L.AadParams()
        assert built_lib.aad_params_default(C.byref(c), kind, 16000) == 0
        mine, _ = fp.to_c()
        for name, _t in L.AadParams._fields_:
            if name in ("custom_fb", "i16_scale", "znorm"):
                continue
            a, b = getattr(c, name), getattr(mine, name)
            assert a == pytest.approx(b), name


def test_invalid_arguments_are_rejected_without_gpu(built_lib):
    assert built_lib.aad_plan_create(None, 0, None) == -1
    assert built_lib.aad_plan_destroy(None) == 0
    assert built_lib.aad_query(None, 1, 10, None, None, None) == -1
    assert built_lib.aad_plan_launches(None) == -1
    assert built_lib.aad_delta(None, None, 1, 1, 1, 9, 1, None, None) == -1


def test_frontend_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import audioanalysisdetector_b200 as aad
    with pytest.raises(aad.AadError):
        aad.Frontend(aad.FrontendParams.mfcc(16000))
    # the per-file drop-ins keep the reference's error convention: print and return None
    assert aad.extract_mfcc((np.zeros(16000, np.float32), 16000)) is None


def test_missing_library_is_an_error(monkeypatch, tmp_path):
    from audioanalysisdetector_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.AadError, match="no CPU fallback"):
        _lib.load()


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "audioanalysisdetector_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_integration_md_struct_matches_binding():
    """The ctypes stub shown to the reference's maintainers must mirror the real aad_params."""
    import os
    import re
    from audioanalysisdetector_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    block = text[text.index("class AadParams(C.Structure)"):text.index("_aad.aad_params_default.argtypes")]
    doc_fields = re.findall(r'\("(\w+)", C\.(?:c_\w+|POINTER\(C\.c_float\))\)', block)
    assert doc_fields == [name for name, _ in _lib.AadParams._fields_]
    header = open(os.path.join(root, "include", "aad.h")).read()
    struct = header[header.index("typedef struct aad_params {"):header.index("} aad_params;")]
    hdr_fields = re.findall(r"^\s*(?:const\s+)?(?:int32_t|float)\s*\*?\s*(\w+);", struct, flags=re.M)
    assert hdr_fields == doc_fields


def test_bind_to_gpu_numa_is_harmless_without_nvml():
    import os
    from audioanalysisdetector_b200 import sharding
    before = os.sched_getaffinity(0)
    cpus = sharding.bind_to_gpu_numa(0)
    assert cpus is None or set(cpus) <= set(before)
    os.sched_setaffinity(0, before)


def test_new_entry_points_validate_their_arguments_before_touching_the_gpu(built_lib):
    """Argument checks of the entry points added around the path (no device work is reached)."""
    from audioanalysisdetector_b200 import _lib as L
    INV = -1
    assert built_lib.aad_extract_indexed(None, None, 0, None, None, 1, 10, None, 0, 1, None, None, None, 0, None) == INV
    assert built_lib.aad_extract_pair(None, None, None, 0, 0, None, None, 1, 10, None, 0, None, 0, 1, None, None, None, 0,
                                      None, 0, None) == INV
    assert built_lib.aad_db_reference(None, 0, 64, None, None, 1, 64, 63, 1, 80.0, None) == INV
    assert built_lib.aad_scaler_accumulate(None, 10, 13, 13, None, None) == INV
    assert built_lib.aad_scaler_apply(None, 10, 13, 13, None, None, None) == INV
    assert built_lib.aad_scaler_accumulate_ragged(None, 4, 10, 13, 13, None, None, None, None) == INV
    w = L.AadDetectorWeights()
    h = C.c_void_p()
    assert built_lib.aad_detector_create(C.byref(w), 0, C.byref(h)) == INV            # struct_size not set
    w.struct_size = C.sizeof(L.AadDetectorWeights)
    w.feature_dim = 13
    assert built_lib.aad_detector_create(C.byref(w), 0, C.byref(h)) == INV            # null weight pointers
    assert built_lib.aad_detector_query(None, 8, None) == INV
    assert built_lib.aad_detector_forward(None, None, 0, 63, 8, None, None, 0, None) == INV
    assert built_lib.aad_detector_destroy(None) == 0
    assert b"paired" in built_lib.aad_strerror(-7)


def test_header_is_plain_c_and_a_c_program_links_against_the_library(built_lib, tmp_path):
    """The boundary is a C ABI: include/aad.h must compile as C99 on its own, and a C program (no C++, no Python) must
    link against libaad_b200.so and get the reference defaults of every plan kind without touching a GPU."""
    import shutil
    import subprocess
    from audioanalysisdetector_b200 import _lib
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    inc = os.path.join(root, "include")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c",
                    os.path.join(inc, "aad.h")], check=True)
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "aad.h"
int main(void) {
  aad_params p;
  int kinds[4] = {AAD_KIND_LOGMEL, AAD_KIND_MFCC, AAD_KIND_LFCC, AAD_KIND_GTCC};
  for (int i = 0; i < 4; ++i) {
    if (aad_params_default(&p, kinds[i], 16000) != AAD_OK) return 1;
    if (p.struct_size != (int)sizeof(aad_params) || p.kind != kinds[i]) return 2;
    printf("%d %d %d %d %d\n", p.kind, p.n_fft, p.hop_length, p.n_filt, p.n_ceps);
  }
  if (aad_params_default(&p, 99, 16000) != AAD_ERR_INVALID_ARG) return 3;
  if (aad_plan_create(NULL, 0, NULL) != AAD_ERR_INVALID_ARG) return 4;
  if (aad_flac_info((const unsigned char*)"nope", 4, NULL) != AAD_ERR_INVALID_ARG) return 5;
  if (strcmp(aad_strerror(AAD_ERR_FORMAT), "malformed or unsupported audio stream") != 0) return 6;
  return aad_version() == AAD_VERSION ? 0 : 7;
}
''')
    exe = tmp_path / "abi"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", inc, str(src), "-o", str(exe), "-L", libdir,
                    "-l:libaad_b200.so", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    assert out[0].split() == ["0", "2048", "512", "64", "0"]        # extract_mel_spectrogram
    assert out[1].split() == ["1", "2048", "512", "128", "13"]      # extract_mfcc
    assert out[2].split() == ["2", "512", "160", "24", "13"]        # extract_lfcc
    assert out[3].split() == ["3", "512", "160", "40", "13"]        # extract_gtcc
