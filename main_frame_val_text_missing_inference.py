#!/usr/bin/env python
"""Drop-in for the reference inference CLI (reference main_frame_val_text_missing_inference.py:247-435):
two no-grad passes per batch, predictions + 4 embeddings per pass.  See sdumc_b200/cli.py."""
from sdumc_b200.cli import main_inference

if __name__ == '__main__':
    main_inference()
