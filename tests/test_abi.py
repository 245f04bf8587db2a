"""CPU tests of the drop-in boundary: libpgs.so loads, exports every symbol include/pgs.h
declares, and refuses to run without a CUDA device (there is no CPU fallback)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_functions():
    text = (ROOT / "include" / "pgs.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pgs_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    names = declared_functions()
    for must in ("pgs_create", "pgs_destroy", "pgs_set_tree", "pgs_set_genes", "pgs_sweep",
                 "pgs_mismatch", "pgs_write_outputs", "pgs_copy_packed", "pgs_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from pangenomesim_b200 import load_library
    lib = load_library()
    for name in declared_functions():
        assert hasattr(lib, name), f"libpgs.so does not export {name}"
    assert lib.pgs_abi_version() == 3


def test_signatures_are_plain_c():
    text = (ROOT / "include" / "pgs.h").read_text()
    assert "torch" not in text and "std::" not in text and "extern \"C\"" in text


def test_genome_name_matches_reference_and_survives_1000_genomes():
    from pangenomesim_b200.capi import genome_name
    assert genome_name(-1) == "genomeref"       # reference util.cxx:17-18
    assert genome_name(0) == "genome001"
    assert genome_name(98) == "genome099"
    assert genome_name(998) == "genome999"
    assert genome_name(999) == "genome1000"     # the reference throws std::length_error here
    assert genome_name(49999) == "genome50000"


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from pangenomesim_b200 import Engine, PgsError
    with pytest.raises(PgsError) as ex:
        Engine(1, 100)
    assert ex.value.code == -6 and "no CPU path" in ex.value.message


def test_product_does_not_reference_the_oracle():
    """the product path must not import, link or call anything under oracle/"""
    for path in list((ROOT / "pangenomesim_b200").rglob("*")):
        if path.suffix in (".py", ".cu", ".cuh", ".cpp", ".h", ".cxx"):
            text = path.read_text()
            assert "oracle_" not in text and "liboracle" not in text and "ref_model" not in text, path
    so = ROOT / "pangenomesim_b200" / "lib" / "libpgs.so"
    assert b"liboracle" not in so.read_bytes()


@pytest.mark.parametrize("sites", [1, 3, 4, 60, 63, 64, 65, 68, 100, 124, 127, 128, 129, 132, 188, 192, 196, 252, 1000, 1001, 1500, 2000, 4096, 4099, 4100])
def test_host_expander_matches_the_oracle(sites):
    """the base expander of PGS_OUT_HOST_EXPAND (SIMD where the CPU has it) against the oracle's
    unpack_ascii and a numpy restatement, every tail length"""
    import numpy as np
    from oracle import oracle_py
    from pangenomesim_b200.capi import expand_bases
    rng = np.random.default_rng(sites)
    words = rng.integers(0, 2**32, size=(sites + 15) // 16 + 3, dtype=np.uint64).astype(np.uint32)
    got, kind = expand_bases(words, sites)
    assert kind in ("avx2", "ssse3", "scalar")
    codes = (words[:, None] >> (2 * np.arange(16, dtype=np.uint32))) & 3
    want = bytes(b"ACGT"[c] for c in codes.reshape(-1)[:sites])
    assert got == want
    padded = np.zeros(4 * ((sites + 63) // 64), np.uint32)
    padded[:len(words)] = words[:len(padded)]
    assert got == oracle_py.unpack_ascii(padded, sites)


def test_plan_only_knob_plans_and_computes_nothing(monkeypatch):
    """PGS_PLAN_ONLY (development knob): an engine without a device runs the host-side planner and
    fails at the first device allocation; nothing can be swept, counted or written with it"""
    import numpy as np
    import treegen
    from pangenomesim_b200 import Engine, PgsError
    monkeypatch.setenv("PGS_PLAN_ONLY", "1")
    parent, rate, leaf = treegen.coalescent(50, seed=3, mut_rate=0.1)
    genes = treegen.gene_table(parent, leaf, n_core=5, n_acc=20, seed=2)
    with Engine(1, 300) as e:
        e.set_tree(parent, rate, leaf)
        with pytest.raises(PgsError) as ex:
            e.set_genes(*genes)
        assert ex.value.code == -6 and "plan-only" in ex.value.message
        with pytest.raises(PgsError) as ex:  # a carrier that is not below its gene's origin is still caught by the planner
            e.set_genes([7], [0], [0, 1], [1])
        assert ex.value.code == -1
        for call in (e.sweep, e.mismatch, lambda: e.write_outputs(None, 8 | 32, [], [])):
            with pytest.raises(PgsError) as ex:
                call()
            assert ex.value.code == -2
        assert e.stats()["n_nodes"] == len(parent) and e.stats()["rows_total"] == 0
    del np



@pytest.mark.parametrize("env", [{}, {"PGS_WALK_PIECE": "41"}, {"PGS_WALK_UNITS": "40", "PGS_WALK_CARVE": "160"},
                                 {"PGS_WALK_UNITS": "1", "PGS_WALK_PIECE": "1"}, {"PGS_WALK_CARVE": "100000"}])
def test_planner_cuts_a_big_tree_on_the_cpu(monkeypatch, env):
    """the dense walk's planner on config 4's tree size and an eighth of its genes (units, carved units,
    bottom-up pieces, every knob at its extreme) runs through on a machine without a device: plan-only
    engines fail at the first device allocation, i.e. after the whole plan has been built"""
    import treegen
    from pangenomesim_b200 import Engine, PgsError
    monkeypatch.setenv("PGS_PLAN_ONLY", "1")
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    n = 10000
    parent, rate, leaf = treegen.coalescent(n, seed=1, mut_rate=0.01)
    with Engine(1, 1000) as e:
        e.set_tree(parent, rate, leaf)
        with pytest.raises(PgsError) as ex:
            e.set_core_genes(625, 2 * n - 2)
        assert ex.value.code == -6 and "plan-only" in ex.value.message


def test_walk_reads_no_row_above_the_dependent_launch_wait():
    """k1_walk_dense is launched as a programmatic dependent of k1_top_masks: nothing that grid writes (origin
    rows, path masks -- the 256-bit row loads) may be read above griddepcontrol.wait.  ptxas moves ld.global.nc
    loads above the wait (SASS: LDG...CONSTANT before ACQBULK), which let a unit start from a stale origin row
    when k1_top_masks was slow to start (seen as rare wrong core genes with several processes on one GPU):
    the prologue's row loads are coherent loads and must stay below ACQBULK in every variant."""
    import shutil
    import subprocess
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(exe).exists():
        pytest.skip("cuobjdump not available")
    so = ROOT / "pangenomesim_b200" / "lib" / "libpgs.so"
    sass = subprocess.run([exe, "-sass", str(so)], capture_output=True, text=True, check=True).stdout
    seen = 0
    name, waited = None, False
    for line in sass.splitlines():
        if "Function :" in line:
            name = line.split("Function :")[1].strip() if "k1_walk_dense" in line else None
            waited = False
            if name:
                seen += 1
            continue
        if not name:
            continue
        if "ACQBULK" in line:
            waited = True
        elif "LDG" in line and ".256" in line:
            assert waited, f"{name}: a row is loaded above griddepcontrol.wait: {line.strip()}"
            assert "CONSTANT" not in line or waited
    assert seen >= 4  # the four instantiations of the walk
