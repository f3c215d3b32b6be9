"""CPU-side checks of the C-ABI boundary: the library builds for sm_100a, loads without a GPU and exports
every symbol include/pianobart_b200.h declares.  No compute calls here."""
import ctypes
import os
import re

from conftest import ROOT


def declared_symbols():
    with open(os.path.join(ROOT, 'include', 'pianobart_b200.h')) as f:
        src = f.read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(pb_[a-z0-9_]+)\s*\(', src)))


def test_library_builds_loads_and_exports_header_symbols():
    from pianobart_b200 import build as B
    lib_path = B.build(verbose=False)
    lib = ctypes.CDLL(lib_path)
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), 'missing export: ' + s
    lib.pb_last_error.restype = ctypes.c_char_p
    assert lib.pb_version() >= 100


def test_sass_contains_tcgen05_and_tma():
    """The GEMM must be a tcgen05/TMA kernel (UTCHMMA / UTMALDG / LDTM in SASS), not an mma.sync port."""
    import subprocess
    from pianobart_b200 import build as B
    obj = os.path.join(B.OBJ, 'gemm_tc.o')
    if not os.path.exists(obj):
        B.build(verbose=False)
    out = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
    assert 'UTCHMMA' in out and 'UTMALDG' in out and 'LDTM' in out
    assert 'HMMA.16816' not in out


def test_no_oracle_import_in_product():
    """The product package must never import the oracle (test infrastructure)."""
    pkg = os.path.join(ROOT, 'pianobart_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            with open(os.path.join(pkg, fn)) as f:
                src = f.read()
            assert 'import oracle' not in src and 'from oracle' not in src, fn


def test_product_fails_loudly_without_gpu():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from pianobart_b200 import _lib
    from pianobart_b200.modules import BartConfig, PianoBart
    from pianobart_b200.vocab import build_octuple_vocab
    e2w, w2e = build_octuple_vocab()
    pb = PianoBart(BartConfig(max_position_embeddings=16, d_model=32, encoder_layers=1, decoder_layers=1,
                              encoder_ffn_dim=32, decoder_ffn_dim=32, encoder_attention_heads=2,
                              decoder_attention_heads=2, vocab_size=8), e2w, w2e)
    ids = torch.zeros(1, 16, 8, dtype=torch.long)
    with pytest.raises(_lib.PBError):
        pb(ids, ids)
