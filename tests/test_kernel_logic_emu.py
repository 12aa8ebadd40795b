"""CPU-side kernel-logic tests: the bodies of the `-m gpu` parity tests, run against
tests/emu/libcm_b200_emu.so — the product's own kernel and host sources (csrc/*.cu) compiled by g++
over a test double of the CUDA runtime that executes a block's threads as fibers. They check the
LOGIC of grouping, scans, packing, the commit protocol over the loopback transport, item building
and both scatter-add kernels against the oracle on a machine without a GPU; timing, memory-model
and NVLink behaviour are what the `-m gpu` suite is for. The emulated library is test
infrastructure: nothing in the product loads it (see tests/emu/include/cuda_runtime.h)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")

GROUPS = [
    "test_single_rank_slices",
    "test_both_scatter_add_kernels",
    "test_single_rank_threshold_flush_and_unsorted_sweep",
    "test_k0_grid_2x2_slices",
    "test_both_exchange_modes",
    "test_nonsquare_grids_wire_bit_exact",
    "test_single_rank_wire_bit_exact",
    "test_edge_cases_duplicates_unsorted_trans_ld",
    "test_elemental_updates",
    "test_symmetric_update_and_fill_lower",
    "test_reset_and_multiple_rounds",
    "test_deterministic_mode_bit_exact_run_to_run",
    "test_cuda_matches_reference_golden",
    "test_accessors_descriptor_host_mirror_and_knobs",
    "test_empty_ragged_and_lopsided_inputs",
    "test_contract_violations_fail_loudly",
    "test_save_load_reference_file_format",
    "test_multi_rank_threshold_flush",
    "test_flush_call_and_staging_limit_several_ranks",
    "test_scatter_add_choice_follows_row_contiguity",
    "device_api_single_rank",
    "pipelined_commit_chunks",
    "panels_wire_and_blocks",
    "panels_unsorted_masked_and_mixed_sources",
    "deferred_apply",
    "deferred_apply_api",
    "two_matrices_and_batch_api",
    "deterministic_mode_under_variants",
    "k1_full_size_properties",
    "bench_step_helpers",
]


@pytest.fixture(scope="module")
def emu_lib():
    subprocess.check_call(["make", "-C", EMU, "-j4"], stdout=subprocess.DEVNULL)
    return os.path.join(EMU, "libcm_b200_emu.so")


def test_group_list_is_complete(emu_lib):
    p = subprocess.run([sys.executable, os.path.join(EMU, "run_cases.py"), "--list"], capture_output=True, text=True,
                       timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    assert sorted(p.stdout.split()) == sorted(GROUPS)


@pytest.mark.parametrize("group", GROUPS)
def test_kernel_logic(emu_lib, group):
    p = subprocess.run([sys.executable, os.path.join(EMU, "run_cases.py"), group], capture_output=True, text=True,
                       timeout=900)
    assert p.returncode == 0 and "fails: 0" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]


def test_cpp_reference_suite_port_on_the_test_double(emu_lib):
    """cpp_tests/cm_test.cpp — the reference's own four scenarios through the C++ drop-in headers,
    T in {int, float, double}, 4 ranks as threads — linked against the kernel-logic test library"""
    p = subprocess.run([os.path.join(EMU, "cm_test_emu"), "2"], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and "PASS" in p.stdout, p.stdout[-2000:] + p.stderr[-3000:]


def test_fuzz_scenarios(emu_lib):
    """a bounded slice of tests/emu/fuzz.py: random geometries / update shapes / index styles / knobs
    (exchange modes, column panels, threshold flushes on several ranks, apply policies, both scatter-add kernels) against the oracle"""
    p = subprocess.run([sys.executable, os.path.join(EMU, "fuzz.py"), "--seed", "0", "--count", "80"], capture_output=True,
                       text=True, timeout=900)
    assert p.returncode == 0 and "fails: 0" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]


def test_product_library_is_not_the_emulator():
    """the product path never loads the test double: libcm_b200.so is an nvcc build (it embeds an
    sm_100a cubin), the package only knows that path, and nothing outside tests/ mentions CM_EMU
    other than the two macros that keep the kernel sources host-compilable"""
    import cm_b200 as cm
    assert cm.LIB_PATH.endswith(os.path.join("convergent-matrix_b200", "libcm_b200.so"))
    with open(os.path.join(ROOT, "convergent-matrix_b200", "__init__.py")) as f:
        assert "emu" not in f.read()
    with open(os.path.join(ROOT, "convergent-matrix_b200", "csrc", "Makefile")) as f:
        assert "CM_EMU" not in f.read()
    if os.path.exists(cm.LIB_PATH):  # no fiber scheduler, no runtime stubs inside the product library
        blob = open(cm.LIB_PATH, "rb").read()
        assert b"cmemu" not in blob and b"cm emu" not in blob
