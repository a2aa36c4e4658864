#!/usr/bin/env python
"""Small end-to-end run for compute-sanitizer (memcheck / initcheck / racecheck): BN and SSM Generators, every kernel family."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import infinite_texture_gans_b200 as itg
from oracle import itg_oracle as O
from common import make_generator
for kw, th, tw in [
    (dict(z_dim=32, G_ch=16, n_layers_G=5, attention=True, leak=0.02, type_norm="BN", outer_padding="replicate"), 3, 4),
    (dict(z_dim=16, G_ch=8, n_layers_G=4, attention=True, leak=0.02, type_norm="SSM", outer_padding="constant"), 3, 3),
    (dict(z_dim=128, G_ch=52, n_layers_G=6, attention=True, leak=0.02, type_norm="BN", outer_padding="replicate"), 2, 2),
]:
    ocfg = O.GenCfg(**kw); sd = O.make_state_dict(ocfg, seed=3, stress=True)
    z, maps = O.make_noise(ocfg, th, tw, seed=4)
    with torch.no_grad(): ref = O.forward_merged(sd, ocfg, z, maps)
    for prec in ("fp16", "fp32"):
        net = make_generator(kw, sd, prec, "cuda")
        img = itg.utils.generate_full_grid(net, z, maps).cpu()
        print(kw["type_norm"], kw["n_layers_G"], prec, "max-abs %.3e" % (img - ref).abs().max().item(), flush=True)
    if kw["n_layers_G"] == 4:
        net = make_generator(kw, sd, "fp16", "cuda")
        P = ocfg.patch_px
        seq = itg.utils.sample_from_gen_PatchByPatch_test(net, z_dim=kw["z_dim"], output_resolution_height=th * P, output_resolution_width=tw * P,
                                                          schedule="sequential", noise=(z, maps))
        print("sequential ok", tuple(seq.shape), flush=True)
print("done")
