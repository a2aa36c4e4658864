#!/usr/bin/env python
"""Drop-in for the reference's test_sample.py (same flags): `python test_sample.py --model_path X.pth
--output_resolution_height H --output_resolution_width W` -- runs the patch-by-patch Generator on the B200 path."""
from infinite_texture_gans_b200.cli import main

if __name__ == "__main__":
    main()
