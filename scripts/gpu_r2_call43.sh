#!/bin/bash
export UMB_LIB_PATH=$PWD/umbrella_b200/libumbrella_b200.tc.so
timeout 150 python scripts/one_attn.py 8 1 128 4100 2>&1 | tail -2
timeout 150 python scripts/one_attn.py 64 8 128 9000 2>&1 | tail -2
