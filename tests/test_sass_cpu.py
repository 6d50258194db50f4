"""What the shipped library actually contains, checked on the SASS of the in-tree .so (cuobjdump; no GPU needed): the
kernels DESIGN.md says run on 5th-generation tensor cores carry tcgen05 MMAs (UTCHMMA / UTCIMMA), TMEM loads (LDTM) and TMA
tile loads (UTMALDG); the CTA-pair GEMM carries the .2CTA forms; the decode kernels carry what the design says they do
(cp.async = LDGSTS ring + warp-level HMMA in the tensor-core GEMV, bulk copies = UBLKCP in the one-kernel decode step) and
NO tcgen05 (the decode-path argument of DESIGN.md section 5).  Mnemonics: /opt/skills/guides/B200_PROFILING.md."""
import collections
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "trtllm-llama_b200", "libtrtllm_llama_b200.so")
WANT = ("UTCHMMA", "UTCIMMA", "UTMALDG", "LDTM", "STTM", "LDGSTS", "UBLKCP", "HMMA", "IMMA", "UTCBAR")


@pytest.fixture(scope="module")
def sass():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(SO) or not os.path.exists(exe):
        pytest.skip("library or cuobjdump not available")
    out = subprocess.run([exe, "-sass", SO], capture_output=True, text=True, timeout=600).stdout
    per = collections.defaultdict(collections.Counter)
    fn = None
    pat = re.compile(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)")
    for line in out.splitlines():
        if "Function :" in line:
            fn = line.split("Function :")[1].strip()
            continue
        m = pat.search(line)
        if fn and m and m.group(1) in WANT:
            per[fn][m.group(1)] += 1
            if ".2CTA" in m.group(2):
                per[fn][m.group(1) + ".2CTA"] += 1
    assert per, "no SASS found: was the library built for sm_100a?"
    return per


def _kernels(per, needle):
    ks = {k: v for k, v in per.items() if needle in k}
    assert ks, f"no kernel matching {needle}"
    return ks


def test_prefill_gemm_is_tcgen05_tma(sass):
    for name, c in _kernels(sass, "gemm_tc_kernel").items():
        assert (c["UTCHMMA"] or c["UTCIMMA"]) and c["UTMALDG"] and c["LDTM"], (name, dict(c))
    for name, c in _kernels(sass, "gemm_tc2_kernel").items():
        assert c["UTCHMMA.2CTA"] or c["UTCIMMA.2CTA"], (name, dict(c))       # cta_group::2
        assert c["UTMALDG"] and c["LDTM"], (name, dict(c))
    # the SmoothQuant instances use the int8 MMA (kind::i8)
    assert any(c["UTCIMMA"] for c in _kernels(sass, "gemm_tc_kernel").values())
    assert any(c["UTCIMMA.2CTA"] for c in _kernels(sass, "gemm_tc2_kernel").values())


def test_prefill_attention_is_tcgen05_tma(sass):
    for name, c in _kernels(sass, "flash_ctx_tc_kernel").items():
        assert c["UTCHMMA"] and c["UTMALDG"] and c["LDTM"] and c["STTM"], (name, dict(c))   # O rescaled in TMEM: tcgen05.st


def test_decode_kernels_are_what_the_design_says(sass):
    for name, c in _kernels(sass, "gemv_mma_kernel").items():
        assert c["LDGSTS"] and (c["HMMA"] or c["IMMA"]), (name, dict(c))     # cp.async weight ring + warp-level MMA
        assert not (c["UTCHMMA"] or c["UTCIMMA"]), (name, dict(c))
    for name, c in _kernels(sass, "decode_step_kernel").items():
        assert c["UBLKCP"], (name, dict(c))                                   # cp.async.bulk weight ring
        assert not (c["UTCHMMA"] or c["UTCIMMA"]), (name, dict(c))
    for name, c in _kernels(sass, "mmha_decode_kernel").items():
        assert not (c["UTCHMMA"] or c["UTCIMMA"]), (name, dict(c))
    assert any(c["HMMA"] for c in _kernels(sass, "mmha_decode_kernel").values())      # the int8-cache tensor-core loops
