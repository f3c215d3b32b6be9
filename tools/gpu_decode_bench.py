"""Decode-only benchmark (BASELINE.json configs[2]): default model, encoder prompt 1024, teacher-forced valid tokens.
usage: python tools/gpu_decode_bench.py [B ...]   (default: 1 64)"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import bench
    from pianobart_b200.modules import BartConfig, PianoBart, PianoBartLM
    from pianobart_b200.vocab import build_octuple_vocab
    c = bench.default_cfg()
    torch.manual_seed(2023)
    e2w, w2e = build_octuple_vocab()
    bc = BartConfig(max_position_embeddings=c['max_pos'], d_model=c['d_model'], encoder_layers=c['layers'],
                    decoder_layers=c['layers'], encoder_ffn_dim=c['ffn'], decoder_ffn_dim=c['ffn'],
                    encoder_attention_heads=c['heads'], decoder_attention_heads=c['heads'])
    dev = torch.device('cuda', 0)
    pb = PianoBart(bc, e2w, w2e, dtype='bf16')
    lm = PianoBartLM(pb).to(dev)
    lm.eval()
    peaks, _ = bench.load_peaks()
    batches = tuple(int(x) for x in sys.argv[1:]) or (1, 64)
    print(json.dumps(bench.decode_bench(lm, pb, dev, peaks, batches=batches)))


if __name__ == '__main__':
    main()
