"""Runs the bodies of the `-m gpu` parity tests (tests/test_gpu_parity.py) against the kernel-logic
TEST library tests/emu/libcm_b200_emu.so — the product's kernel and host sources compiled by g++
over a test double of the CUDA runtime (tests/emu/include/cuda_runtime.h) — so that packing, the
commit protocol, item building and both scatter-add kernels are checked against the oracle on a
machine without a GPU. Test infrastructure only: the product library is libcm_b200.so (nvcc,
sm_100a) and this file never touches it.

usage: python tests/emu/run_cases.py [test function name ...]     (no names = every case)
"""
import os
import pathlib
import sys
import tempfile
import time
import traceback

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cm_b200 as cm  # noqa: E402

cm.LIB_PATH = os.environ.get("CM_EMU_LIB", os.path.join(HERE, "libcm_b200_emu.so"))  # e.g. a sanitizer build  # explicit: the tests' library, not the product's

import test_gpu_parity as T  # noqa: E402
import test_gpu_variants as V  # noqa: E402
import emu_cases as E  # noqa: E402


def _tmp():
    return pathlib.Path(tempfile.mkdtemp(prefix="cm_emu_"))


CASES = [
    (T, "test_single_rank_slices", [(cm.F64,), (cm.F32,), (cm.I32,)]),
    (T, "test_both_scatter_add_kernels", [(k, g) for k in (1, 2) for g in
                                          [(1, 1, 2048, 2048, 2048), (2, 2, 2000, 1000, 1024), (1, 1, 20000, 300, 20000)]]),
    (T, "test_single_rank_threshold_flush_and_unsorted_sweep", [()]),
    (T, "test_k0_grid_2x2_slices", [(cm.F64,), (cm.F32,), (cm.I32,)]),
    (T, "test_both_exchange_modes", [(1,), (2,)]),
    (T, "test_nonsquare_grids_wire_bit_exact", [((2, 3, 4, 8, 64, 100, 190),), ((3, 2, 8, 4, 96, 250, 150),),
                                                ((2, 4, 64, 64, 1024, 2000, 3000),), ((1, 2, 64, 64, 1024, 1000, 2000),),
                                                ((2, 1, 64, 64, 1024, 2000, 1000),)]),
    (T, "test_single_rank_wire_bit_exact", [()]),
    (T, "test_edge_cases_duplicates_unsorted_trans_ld", [()]),
    (T, "test_elemental_updates", [(1,), (4,)]),
    (T, "test_symmetric_update_and_fill_lower", [(1,), (4,)]),
    (T, "test_reset_and_multiple_rounds", [()]),
    (T, "test_deterministic_mode_bit_exact_run_to_run", [()]),
    (T, "test_cuda_matches_reference_golden", [(n,) for n in sorted(T.SCENARIOS)]),
    (T, "test_accessors_descriptor_host_mirror_and_knobs", [()]),
    (T, "test_empty_ragged_and_lopsided_inputs", [()]),
    (T, "test_contract_violations_fail_loudly", [()]),
    (T, "test_save_load_reference_file_format", [(1, None), (4, None)]),
    (T, "test_multi_rank_threshold_flush", [(2,), (1,)]),
    (T, "test_flush_call_and_staging_limit_several_ranks", [()]),
    (T, "test_scatter_add_choice_follows_row_contiguity", [()]),
] + [(E, name, params) for name, params in E.CASES]

NAMES = [name for _, name, _ in CASES]


def main(argv):
    if argv == ["--list"]:
        print("\n".join(NAMES))
        return 0
    only = set(argv)
    unknown = only - set(NAMES)
    if unknown:
        print("unknown case(s):", sorted(unknown))
        return 2
    fails = 0
    for mod, name, params in CASES:
        if only and name not in only:
            continue
        for p in params:
            if name == "test_save_load_reference_file_format":
                p = (p[0], _tmp())
            t = time.time()
            try:
                getattr(mod, name)(*p)
                print(f"PASS {name}{p if len(repr(p)) < 70 else '(...)'} {time.time() - t:.2f}s", flush=True)
            except Exception as e:  # noqa: BLE001
                fails += 1
                print(f"FAIL {name}{p}: {type(e).__name__}: {str(e)[:400]}", flush=True)
                traceback.print_exc(limit=6)
    print("fails:", fails, flush=True)
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
