"""A few seconds per op of tools/stress_emu.py (random alphabets, run structures, zlib levels / strategies, damaged inputs, ragged
layouts) on the CPU emulator against the oracle.  The tool itself runs as long as it is given: `python tools/stress_emu.py 600`."""
import random

import pytest

from tools import stress_emu


@pytest.mark.parametrize("op", ["dc", "flate", "ari", "lz4", "bwt", "mtf_rle", "pipeline", "zlib_adler", "encoders"])
def test_stress_emu(emu_ctx, op, monkeypatch):
    monkeypatch.setenv("RCZ_LZ4_CHUNK_BYTES", "30000")
    monkeypatch.setenv("RCZ_HOST_CHUNK_BYTES", "20000")
    fn = getattr(stress_emu, "stress_" + op)
    assert fn(emu_ctx, random.Random(2024), 3.0) > 0
