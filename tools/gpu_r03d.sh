#!/bin/bash
# forward attention A/B: speculative first chunk (a -DSMBV_ATTN_SPEC build, measured neutral and removed again) vs base, interleaved runs on one box
mkdir -p gpurun_out
for r in 1 2 3; do
  python tools/attn_fwd_ab.py base
  SMBV_LIB=$PWD/smb-vision_b200/lib/libsmbv_b200_spec.so python tools/attn_fwd_ab.py spec
done 2>&1 | tee gpurun_out/attn_fwd_spec_ab.log
