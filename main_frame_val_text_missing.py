#!/usr/bin/env python
"""Drop-in for the reference training CLI (same flags and stdout lines; reference
main_frame_val_text_missing.py:209-417) on the B200-native hot path.  See sdumc_b200/cli.py."""
from sdumc_b200.cli import main_train

if __name__ == '__main__':
    main_train()
