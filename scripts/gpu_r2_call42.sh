#!/bin/bash
export UMB_LIB_PATH=$PWD/umbrella_b200/libumbrella_b200.tcdbg.so
timeout 200 python scripts/one_attn_dbg.py 2>&1 | tail -22
