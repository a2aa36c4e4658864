"""`test_sample.py` of the reference (test_sample.py:11-79) on the B200 path: same flags, same checkpoint format
({'args': Namespace, 'netG_state_dict': ...}, optional 'module.' prefixes), same output convention (img*0.5+0.5 saved
next to the checkpoint).  The extra flags default to what reproduces the reference: `--schedule auto` runs the shipped sequential
3x3 schedule whenever the checkpoint's attention block contributes (gamma != 0) and the equivalent one-shot pass otherwise."""
from __future__ import annotations

import argparse
import os
from collections import OrderedDict

import torch


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser()
    # the reference's flags (test_sample.py:14-18)
    p.add_argument("--output_resolution_height", type=int, default=384, help="output_resolution_height")
    p.add_argument("--output_resolution_width", type=int, default=384, help="output_resolution_width")
    p.add_argument("--output_name", type=str, default="241_generated.jpg", help="name of the generated image")
    p.add_argument("--model_path", type=str, default="results/241_lp_bn_outerpadRepl/300__ema.pth", help="path of the generator network")
    p.add_argument("--tiles", default=False, action="store_true", help="use tiling of the input (only read on the non-local path, test_sample.py:70-73)")
    # additions
    p.add_argument("--precision", default="fp16", choices=["fp16", "fp32"], help="operand precision of the CUDA path (16-bit tensor-core mode / fp32 exact mode)")
    p.add_argument("--schedule", default="auto", choices=["auto", "oneshot", "sequential"],
                   help="oneshot: whole patch grid in one device-resident pass; sequential: the shipped 3x3 sub-image schedule; "
                        "auto: sequential iff the attention block contributes (gamma != 0), i.e. whenever the two differ")
    p.add_argument("--unsafe-load", action="store_true",
                   help="unpickle the checkpoint without torch.load's weights_only guard (only for checkpoints you trust)")
    p.add_argument("--seed", type=int, default=None, help="torch.manual_seed before drawing z (the reference has no seed flag)")
    return p


def load_G(state_dict_G, netG):
    """test_sample.py:32-41: strip 'module.' (nn.DataParallel checkpoints), load, eval."""
    new_sd = OrderedDict()
    for k, v in state_dict_G.items():
        new_sd[k.replace("module.", "") if "module" in k else k] = v
    netG.load_state_dict(new_sd)
    netG.eval()
    return netG


def main(argv=None) -> str:
    from . import generators, utils
    a = build_parser().parse_args(argv)
    if not torch.cuda.is_available():
        raise SystemExit("infinite_texture_gans_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    device = torch.device("cuda:0")
    folder, _ = os.path.split(a.model_path)
    # checkpoints pickle the training Namespace next to the tensors (train.py:207): allow-list it and keep the weights_only guard
    torch.serialization.add_safe_globals([argparse.Namespace])
    ckpt = torch.load(a.model_path, map_location="cpu", weights_only=not a.unsafe_load)
    args, sd = ckpt["args"], ckpt["netG_state_dict"]
    netG = generators.ResidualPatchGenerator(
        z_dim=args.z_dim, G_ch=args.G_ch, base_res=args.base_res, n_layers_G=args.n_layers_G, attention=args.attention,
        img_ch=args.img_ch, leak=args.leak_G, SN=False, type_norm=args.type_norm_G, map_dim=1, padding_mode=args.padding_mode,
        outer_padding=args.outer_padding, num_patches_h=3, num_patches_w=3, padding_size=1, conv_reduction=2,
        precision=a.precision)
    netG = load_G(sd, netG).to(device)
    print(args)
    if a.seed is not None:
        torch.manual_seed(a.seed)
    with torch.no_grad():
        if args.padding_mode != "local":                 # test_sample.py:70-73: the non-local Generator, optionally tiled
            scale = 2 ** (netG.n_layers_G - 1)
            img = utils.sample_from_gen(netG, z_dim=args.z_dim, base_res=a.output_resolution_height // scale, num_images=1, tiles=a.tiles,
                                        device=device)
        else:
            if a.tiles:
                print("--tiles only applies to --padding_mode zeros checkpoints (test_sample.py:70-73); ignored for local padding")
            img = utils.sample_from_gen_PatchByPatch_test(
                netG, z_dim=args.z_dim, base_res=args.base_res, num_images=1, output_resolution_height=a.output_resolution_height,
                output_resolution_width=a.output_resolution_width, device=device, schedule=a.schedule, return_on_device=True)
        # test_sample.py:78 `save_image(img * 0.5 + 0.5, path)`: torchvision quantises with mul(255).add_(0.5).clamp_(0, 255).to(uint8) and
        # hands the (H, W, C) bytes to PIL.  The same bytes are produced on the device (a quarter of the PCIe traffic of the fp32 image).
        arr = utils.image_to_uint8(img).cpu().numpy()
    path = os.path.join(folder, a.output_name)
    print("The image is saved as:", path)
    from PIL import Image
    if arr.shape[2] == 1:                    # save_image -> make_grid replicates a single channel to RGB (torchvision/utils.py)
        arr = arr.repeat(3, axis=2)
    Image.fromarray(arr).save(path)
    return path


if __name__ == "__main__":
    main()
