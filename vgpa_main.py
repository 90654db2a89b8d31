#!/usr/bin/env python
"""Command line of the reference (vgpa_main.py:149-179): --params <json> [--data <csv>]."""
import argparse

from vgpa_b200.simulation import main

if __name__ == "__main__":
    parser = argparse.ArgumentParser(description="VGPA on B200: variational inference for SDEs.")
    parser.add_argument("--params", type=str, help="Input file (.json) with simulation parameters.")
    parser.add_argument("--data", type=str, default=None, help="Input file (.csv) with observations.")
    args = parser.parse_args()
    main(args.params, args.data)
