"""CPU check of the fused tcgen05 kernel SOURCE in the host CTA emulator (tests/emu).

wetts_b200/csrc/fused_rb_kernel.cuh contains no PTX; compiled with -DWETTS_EMULATE every primitive maps to
tests/emu/emu_runtime.h (one OS thread per CUDA thread, lazily delivered bulk copies and MMAs, mbarrier phase
semantics, TMEM lane-quarter rule).  The driver tests/emu/fused_rb_emu.cpp compares the kernel's output with an
fp64 evaluation of ResBlock2 x nrb + MRF mean (decoders.py:205-214, :72-76) and exits non-zero above 2e-5 of
rms.  This exercises index arithmetic, descriptors, packing, the weight ring and every barrier hand-off without
a GPU; the `-m gpu` tests then check the same kernel on hardware through the C ABI."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")


@pytest.fixture(scope="module")
def emu_binary(tmp_path_factory):
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("g++ not available")
    out = str(tmp_path_factory.mktemp("emu") / "fused_rb_emu")
    cmd = [gxx, "-O2", "-std=c++20", "-pthread", "-x", "c++", "-I", EMU, "-I", os.path.join(ROOT, "wetts_b200", "csrc"),
           os.path.join(EMU, "fused_rb_emu.cpp"), "-o", out]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return out


# (C, B, T, grid, nrb, ring slots): ragged last tile, single short tile, more CTAs than items, 1-3 resblocks,
# both widths, both ring sizes (6 slots need a chunk count per item that is a multiple of 6: nrb = 3 or 1 here)
CASES = [(32, 2, 300, 2, 3, 4), (32, 1, 76, 1, 3, 6), (32, 2, 256, 5, 1, 6), (32, 2, 520, 3, 2, 4),
         (64, 2, 300, 2, 3, 6), (64, 1, 640, 2, 3, 4), (64, 3, 320, 4, 2, 4), (32, 2, 1000, 3, 3, 6)]


@pytest.mark.parametrize("case", CASES)
def test_fused_resblock_kernel_in_emulator(emu_binary, case):
    r = subprocess.run([emu_binary] + [str(v) for v in case], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "rel=" in r.stdout


# ---- the emulator must be able to FAIL: mutate the kernel source and expect wrong numbers / a trapped deadlock
MUTATIONS = {
    # epilogue 1 reads TMEM without waiting for the conv1 accumulator barrier -> stale accumulators
    "no_acc1_wait": ("      mbar_wait(bar_acc1, rb_count & 1);", "      // (mutation) no wait"),
    # the producer recycles a ring slot without waiting for the MMAs that still read it -> weights overwritten early
    "no_empty_wait": ("    if (use > 0) mbar_wait(bar_empty + 8 * slot, (use - 1) & 1);", "    // (mutation) no wait"),
    # every tap reads one row too far (index arithmetic defect)
    "tap_shift_off_by_one": ("row0 + m * row_step + tap * dil);", "row0 + m * row_step + tap * dil + 1);"),
}


@pytest.mark.parametrize("name", sorted(MUTATIONS))
def test_emulator_detects_broken_kernels(tmp_path, name):
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("g++ not available")
    csrc = os.path.join(ROOT, "wetts_b200", "csrc")
    for f in ("fused_rb_kernel.cuh", "fused_rb_args.h", "tc_prims.cuh"):
        shutil.copy(os.path.join(csrc, f), tmp_path / f)
    old, new = MUTATIONS[name]
    src = (tmp_path / "fused_rb_kernel.cuh").read_text()
    assert src.count(old) == 1, f"mutation anchor for {name} not found exactly once"
    (tmp_path / "fused_rb_kernel.cuh").write_text(src.replace(old, new))
    out = str(tmp_path / "emu_mut")
    subprocess.run([gxx, "-O2", "-std=c++20", "-pthread", "-x", "c++", "-I", EMU, "-I", str(tmp_path),
                    os.path.join(EMU, "fused_rb_emu.cpp"), "-o", out], check=True, capture_output=True, text=True)
    r = subprocess.run([out, "32", "2", "300", "2", "3", "4"], capture_output=True, text=True, timeout=600)
    assert r.returncode != 0, "the emulator accepted a kernel with a known defect:\n" + r.stdout
