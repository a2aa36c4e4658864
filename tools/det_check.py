import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import infinite_texture_gans_b200 as itg
from oracle import itg_oracle as O
from common import make_generator
kw = dict(z_dim=128, G_ch=52, n_layers_G=6, attention=True, leak=0.02, type_norm="BN", outer_padding="replicate")
ocfg = O.GenCfg(**kw); sd = O.make_state_dict(ocfg, seed=101, stress=True)
z, maps = O.make_noise(ocfg, 2, 4, seed=102)
with torch.no_grad(): ref = O.forward_merged(sd, ocfg, z, maps)
outs = []
for i in range(4):
    net = make_generator(kw, sd, "fp32", "cuda")
    img = itg.utils.generate_full_grid(net, z, maps).cpu().clone()
    outs.append(img)
    print(i, 'err %.3e' % (img - ref).abs().max().item(), 'vs run0 %.3e' % (img - outs[0]).abs().max().item())
d = (outs[-1] - ref).abs()
idx = d.flatten().topk(5).indices
print([(int(i // (d.shape[2]*d.shape[3])), int(i // d.shape[3] % d.shape[2]), int(i % d.shape[3])) for i in idx], d.flatten()[idx])
