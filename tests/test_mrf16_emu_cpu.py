"""CPU check of the f16-split fused MRF kernel SOURCE (wetts_b200/csrc/fused_mrf16_kernel.cuh) in the host CTA emulator.

Same method as tests/test_fused_emu_cpu.py: the kernel contains no PTX, every primitive maps to tests/emu/emu_runtime.h
(kind::f16 MMAs evaluated from the shared-memory descriptors, K = 16, fp16 operands), and the driver
tests/emu/fused_mrf16_emu.cpp compares against an fp64 evaluation of ResBlock1 / ResBlock2 x nrb + the MRF mean
(decoders.py:157-170, :205-214, :72-76); it exits non-zero above 2e-5 of rms or if a skipped tile was written."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")


def _build(out, include_dir):
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("g++ not available")
    subprocess.run([gxx, "-O2", "-std=c++20", "-pthread", "-x", "c++", "-I", EMU, "-I", include_dir,
                    os.path.join(EMU, "fused_mrf16_emu.cpp"), "-o", out], check=True, capture_output=True, text=True)
    return out


@pytest.fixture(scope="module")
def emu_binary(tmp_path_factory):
    return _build(str(tmp_path_factory.mktemp("emu16") / "fused_mrf16_emu"), os.path.join(ROOT, "wetts_b200", "csrc"))


# (type, C, B, T, grid, nrb, ring, threads, length-aware[, item rows]): ResBlock2 (v3) and ResBlock1 (v1) stages, both widths, both
# ring sizes, ragged last tile, single short tile, more CTAs than items, 1-3 resblocks, the length-aware item list
CASES = [(2, 32, 2, 300, 2, 3, 4, 256, 0), (2, 32, 1, 76, 1, 3, 6, 256, 0), (2, 32, 2, 256, 5, 1, 6, 256, 0),
         (2, 64, 2, 300, 2, 3, 6, 512, 0), (2, 64, 3, 320, 4, 2, 4, 512, 0), (2, 32, 4, 1000, 3, 3, 6, 256, 1),
         (1, 32, 2, 300, 2, 3, 6, 256, 0), (1, 32, 1, 76, 1, 3, 4, 256, 0), (1, 64, 2, 300, 2, 3, 6, 512, 0),
         (1, 32, 4, 700, 3, 3, 6, 256, 1),
         # 256-sample work items (ITEM = 256: three M blocks per conv, two per resblock output, two prefetched staging units)
         (2, 32, 2, 300, 2, 3, 6, 256, 0, 256), (2, 32, 1, 76, 1, 3, 6, 256, 0, 256), (2, 32, 4, 1000, 3, 3, 6, 256, 1, 256),
         (2, 64, 2, 600, 2, 3, 6, 512, 0, 256), (1, 32, 2, 600, 2, 3, 6, 256, 0, 256), (2, 32, 2, 512, 5, 1, 6, 256, 0, 256),
         # 384-sample items (the default of the C = 32 ResBlock2 stage for large launches: four M blocks = 256 TMEM columns)
         (2, 32, 2, 900, 2, 3, 6, 256, 0, 384), (2, 32, 4, 1000, 3, 3, 6, 256, 1, 384), (2, 32, 1, 76, 1, 3, 6, 256, 0, 384),
         (2, 32, 2, 900, 2, 2, 4, 256, 0, 384), (2, 32, 1, 388, 2, 3, 6, 256, 0, 384), (2, 32, 2, 4, 3, 3, 6, 256, 0, 384)]


@pytest.mark.parametrize("case", CASES)
def test_fused_mrf16_kernel_in_emulator(emu_binary, case):
    r = subprocess.run([emu_binary] + [str(v) for v in case], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "rel=" in r.stdout


# ---- the emulator must be able to FAIL on this kernel too
MUTATIONS = {
    # the epilogue reads TMEM without waiting for the accumulator barrier -> stale accumulators
    "no_acc_wait": ("        mbar_wait(bar_acc, conv_count & 1);", "        // (mutation) no wait"),
    # the small-term columns are added without the 2^-11 scale
    "lo_scale_missing": ("vs[8 * g8 + e] * kF16LoInv", "vs[8 * g8 + e]"),
    # ResBlock1's inner conv result lands in the X tile instead of T (the residual input is destroyed)
    "inner_conv_overwrites_x": ("uint8_t* dst_tile = inner ? Tt : Xt;", "uint8_t* dst_tile = Xt;"),
}


@pytest.mark.parametrize("name", sorted(MUTATIONS))
def test_emulator_detects_broken_mrf16_kernels(tmp_path, name):
    csrc = os.path.join(ROOT, "wetts_b200", "csrc")
    for f in ("fused_mrf16_kernel.cuh", "fused_mrf16_args.h", "tc_prims.cuh"):
        shutil.copy(os.path.join(csrc, f), tmp_path / f)
    old, new = MUTATIONS[name]
    src = (tmp_path / "fused_mrf16_kernel.cuh").read_text()
    assert src.count(old) == 1, f"mutation anchor for {name} not found exactly once"
    (tmp_path / "fused_mrf16_kernel.cuh").write_text(src.replace(old, new))
    out = _build(str(tmp_path / "emu_mut"), str(tmp_path))
    args = ["1", "32", "2", "300", "2", "3", "6"] if name == "inner_conv_overwrites_x" else ["2", "32", "2", "300", "2", "3", "4"]
    r = subprocess.run([out] + args, capture_output=True, text=True, timeout=900)
    assert r.returncode != 0, "the emulator accepted a kernel with a known defect:\n" + r.stdout


# ---- tensor-pipe attention kernel (attn_tc_kernel.cuh) in the same emulator
@pytest.fixture(scope="module")
def attn_emu_binary(tmp_path_factory):
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("g++ not available")
    out = str(tmp_path_factory.mktemp("emu_attn") / "attn_tc_emu")
    subprocess.run([gxx, "-O2", "-std=c++20", "-pthread", "-x", "c++", "-I", EMU, "-I", os.path.join(ROOT, "wetts_b200", "csrc"),
                    os.path.join(EMU, "attn_tc_emu.cpp"), "-o", out], check=True, capture_output=True, text=True)
    return out


# (B, T, lengths...): full tile, ragged lengths incl. a nearly empty utterance (invalid query rows), short texts
@pytest.mark.parametrize("shared", [0, 1])
@pytest.mark.parametrize("case", [(2, 128, 128, 80), (3, 100, 100, 37, 1), (1, 64, 64), (2, 7, 7, 3)])
def test_attention_tc_kernel_in_emulator(attn_emu_binary, case, shared):
    """windowed relative-position attention (attentions.py:232-282) on the emulated tensor pipe vs an fp64 evaluation;
    shared = 1: the two-CTAs-per-SM variant (shared memory used twice, O in the TMEM columns of S)"""
    r = subprocess.run([attn_emu_binary] + [str(v) for v in case], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, EMU_ATTN_SHARED=str(shared)))
    assert r.returncode == 0, r.stdout + r.stderr


# ---- pipelined per-layer conv kernel (tc16p_conv_kernel.cuh) in the same emulator
@pytest.fixture(scope="module")
def tc16p_emu_binary(tmp_path_factory):
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("g++ not available")
    out = str(tmp_path_factory.mktemp("emu_tc16p") / "tc16p_emu")
    subprocess.run([gxx, "-O2", "-std=c++20", "-pthread", "-x", "c++", "-I", EMU, "-I", os.path.join(ROOT, "wetts_b200", "csrc"),
                    os.path.join(EMU, "tc16p_emu.cpp"), "-o", out], check=True, capture_output=True, text=True)
    return out


# (Cin, Cout, K, dil, B, T, N, KC, grid, epilogue mode, length-aware): multi-chunk K loop, taps / dilation, several N tiles,
# ragged T and masks, plain+relu / residual / gate epilogues, 2- and 4-slot activation rings, zero-tile items
@pytest.mark.parametrize("case", [(48, 64, 5, 1, 2, 300, 64, 16, 2, 0, 0), (48, 64, 5, 1, 2, 300, 64, 16, 2, 1, 0),
                                  (96, 128, 3, 2, 3, 520, 64, 32, 2, 2, 0), (40, 96, 7, 3, 2, 700, 32, 16, 3, 0, 1),
                                  (192, 128, 1, 1, 2, 260, 128, 64, 1, 1, 0), (32, 64, 3, 12, 1, 1000, 64, 32, 1, 1, 1)])
def test_pipelined_conv_kernel_in_emulator(tc16p_emu_binary, case):
    # default role assignment of the harness and of the launcher: all worker warps stage, then drain; one accumulator set.
    # EMU_TIMEOUT_S: a drain warp legitimately waits for a whole emulated item of MMAs (slow on a loaded machine)
    r = subprocess.run([tc16p_emu_binary] + [str(v) for v in case], capture_output=True, text=True, timeout=2400,
                       env=dict(os.environ, EMU_TIMEOUT_S="900"))
    assert r.returncode == 0, r.stdout + r.stderr


# (NOT covered: the assignments with dedicated drain warps, EMU_ALLWARPS=0 -- with or without the two-accumulator-slot
# experiment they deadlock intermittently in the emulator (one run in ~5 under load, unexplained); they are experiments
# behind environment switches, off by default in the launcher, and measured slower than the default kernel)
@pytest.mark.parametrize("env", [dict(EMU_PINGPONG="1", EMU_ALLWARPS="1"), dict(EMU_NTMINOR="0", EMU_NB="2"), dict(EMU_NB="3")])
def test_pipelined_conv_kernel_role_assignments(tc16p_emu_binary, env):
    """the other role assignments / ring depths of the pipelined kernel on one multi-tile, multi-chunk gate-epilogue case"""
    r = subprocess.run([tc16p_emu_binary] + [str(v) for v in (96, 128, 3, 2, 2, 520, 64, 32, 2, 2, 0)], capture_output=True, text=True,
                       timeout=2400, env=dict(os.environ, EMU_TIMEOUT_S="900", **env))
    assert r.returncode == 0, r.stdout + r.stderr


# ---- row-block-resident per-layer conv kernel (tc16r_conv_kernel.cuh) in the same emulator
@pytest.fixture(scope="module")
def tc16r_emu_binary(tmp_path_factory):
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("no g++")
    out = str(tmp_path_factory.mktemp("emu_tc16r") / "tc16r_emu")
    subprocess.run([cxx, "-O2", "-std=c++20", "-pthread", "-x", "c++", "-I", EMU, "-I", os.path.join(ROOT, "wetts_b200", "csrc"),
                    os.path.join(EMU, "tc16r_emu.cpp"), "-o", out], check=True, capture_output=True, text=True)
    return out


# (Cin, Cout, K, dil, B, T, N, KC, grid, epilogue mode, length-aware): 1..4 N tiles through the two TMEM halves, the shapes of
# a flow in_layer (192 -> 384, k5, gate) and res_skip (1x1, residual), a polyphase upsampler (k2), ragged T, skipped blocks
@pytest.mark.parametrize("case", [(48, 64, 5, 1, 2, 300, 64, 16, 2, 0, 0), (96, 256, 3, 2, 2, 520, 128, 32, 2, 2, 0),
                                  (192, 384, 1, 1, 2, 260, 128, 64, 2, 1, 0), (40, 96, 7, 3, 2, 700, 32, 16, 3, 0, 1),
                                  (64, 256, 3, 1, 2, 500, 64, 32, 3, 1, 1), (192, 384, 5, 1, 2, 400, 128, 16, 2, 2, 0),
                                  (128, 512, 2, 1, 2, 300, 128, 48, 2, 0, 0)])
def test_resident_conv_kernel_in_emulator(tc16r_emu_binary, case):
    r = subprocess.run([tc16r_emu_binary] + [str(v) for v in case], capture_output=True, text=True, timeout=2400,
                       env=dict(os.environ, EMU_TIMEOUT_S="600"))
    assert r.returncode == 0, r.stdout + r.stderr
