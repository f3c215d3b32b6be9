"""Short decode run (no CUDA graph) for an ncu launch list of one decode step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import params as P
from pianobart_b200.generate import Generator
from pianobart_b200.modules import BartConfig, PianoBart, PianoBartLM
from pianobart_b200.vocab import build_octuple_vocab
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
e2w, w2e = build_octuple_vocab()
bc = BartConfig(max_position_embeddings=1024, d_model=1024, encoder_layers=8, decoder_layers=8, encoder_ffn_dim=2048,
                decoder_ffn_dim=2048, encoder_attention_heads=8, decoder_attention_heads=8)
pb = PianoBart(bc, e2w, w2e, dtype='bf16'); lm = PianoBartLM(pb).cuda().eval()
S = 1024
gen = Generator(lm, B, S, S, use_graph=False)
ids = torch.from_numpy(P.synth_ids(B, S, 1)).cuda(); keep = torch.ones(B, S, device='cuda')
forced = torch.from_numpy(P.synth_ids(B, S, 2)).cuda()
gen.start(ids, keep, np.zeros((B, S, 8)), forced)
gen.run_steps(300 if len(sys.argv) > 2 else 6)
torch.cuda.synchronize()
