"""CPU suite for the host layer and the C-ABI boundary (no compute calls: there is no GPU here)."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cdq():
    import chess_deep_q_b200 as m
    return m


@pytest.fixture(scope="module")
def shim_chess():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "pychess_shim"))
    import chess
    return chess


def test_library_builds_and_exports_every_declared_symbol(cdq):
    import importlib.util
    spec = importlib.util.spec_from_file_location("cdq_build", os.path.join(ROOT, "chess-deep-q_b200", "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    so = b.build()
    assert os.path.exists(so)
    header = open(os.path.join(ROOT, "include", "cdq.h")).read()
    declared = set(re.findall(r"\b(cdq_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 14
    L = ctypes.CDLL(so)
    for name in declared:
        assert hasattr(L, name), "libcdq.so does not export %s" % name
    assert declared == set(cdq._lib.SIGNATURES), declared ^ set(cdq._lib.SIGNATURES)
    L.cdq_version.restype = ctypes.c_int
    assert L.cdq_version() == 100
    L.cdq_error_string.restype = ctypes.c_char_p
    assert b"no CPU fallback" in L.cdq_error_string(-2)


def test_sass_uses_blackwell_tensor_and_tma_paths():
    import subprocess
    so = os.path.join(ROOT, "chess-deep-q_b200", "libcdq.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass      # tcgen05.mma
    assert "LDTM" in sass         # tcgen05.ld
    assert "UBLKCP" in sass       # bulk async copy (TMA engine)
    assert "HMMA." not in sass.replace("UTCHMMA", "")   # no legacy mma.sync path


def test_board_dtype_matches_header(cdq, oracle):
    assert cdq.BOARD_DTYPE == oracle.BOARD_DTYPE
    assert cdq.BOARD_DTYPE.itemsize == 72


def test_pack_board_matches_oracle_records(cdq, oracle, shim_chess):
    fens = ["rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1",
            "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1",
            "rnbqkbnr/ppp1pppp/8/8/3pP3/8/PPPP1PPP/RNBQKBNR b KQkq e3 0 2",
            "4k3/8/8/8/8/8/8/R3K2R b KQ - 0 1", "8/8/4k3/8/8/3K4/8/8 w - - 0 1"]
    for f in fens:
        rec = cdq.pack_board(shim_chess.Board(f))
        assert rec.tobytes() == oracle.from_fen(f).tobytes(), f
    recs = cdq.pack_boards([shim_chess.Board(f) for f in fens])
    assert recs.shape == (5,) and cdq.pack_boards([]).shape == (0,)


def test_pack_board_sets_rep3_from_history(cdq, shim_chess):
    b = shim_chess.Board()
    for u in ["g1f3", "g8f6", "f3g1", "f6g8"] * 2:
        b.push([m for m in b.legal_moves if m.uci() == u][0])
    assert int(cdq.pack_board(b)["meta"][0]) & (1 << 16)
    assert not int(cdq.pack_board(shim_chess.Board())["meta"][0]) & (1 << 16)


def test_as_records_accepts_all_forms(cdq, oracle):
    from chess_deep_q_b200 import board_utils as bu
    rec = oracle.playouts(10, seed=1)
    assert np.array_equal(bu.as_records(rec), rec)
    as_u64 = rec.view(np.uint64).reshape(-1, 9)
    assert np.array_equal(bu.as_records(as_u64), rec)
    assert np.array_equal(bu.as_records(as_u64.view(np.int64)), rec)
    with pytest.raises(TypeError):
        bu.as_records(np.zeros((3, 4)))


def test_planes_to_packed_roundtrip_cpu(cdq, oracle):
    import torch
    from chess_deep_q_b200 import board_utils as bu
    rec = oracle.playouts(50, seed=2)
    planes = torch.from_numpy(oracle.encode_planes(rec))
    back = bu.planes_to_packed(planes).numpy().view(np.uint64)
    assert np.array_equal(back[:, :8], rec["bb"])


def test_state_dict_keys_match_reference_checkpoint_layout(cdq):
    sd = cdq.ChessQNetwork().state_dict()
    want = {"conv1.weight": (32, 12, 3, 3), "conv1.bias": (32,), "conv2.weight": (64, 32, 3, 3),
            "conv2.bias": (64,), "fc1.weight": (256, 4096), "fc1.bias": (256,), "fc2.weight": (1, 256),
            "fc2.bias": (1,)}
    assert {k: tuple(v.shape) for k, v in sd.items()} == want
    assert sum(v.numel() for v in sd.values()) == 1071073


def test_compute_fails_loudly_without_gpu(cdq, oracle):
    import torch
    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    with pytest.raises(cdq._lib.CdqError):
        cdq.fast_evaluate_position(oracle.startpos())
    with pytest.raises(cdq._lib.CdqError):
        cdq.ChessQNetwork().forward_packed(oracle.startpos())
    with pytest.raises(cdq._lib.CdqError):
        cdq.DQNAgent()


def test_product_never_imports_the_oracle():
    """The product path must not route through oracle/ (tier rule 3)."""
    pkg = os.path.join(ROOT, "chess-deep-q_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the CPU oracle", "").replace("CPU oracle", ""), f
