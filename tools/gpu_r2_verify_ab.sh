#!/bin/bash
# Per-kernel launch lists of one MNT4-753 chunk verification: current build vs the variant with the round-2-start formulas
TAG=${1:-r2r}
bash tools/gpu_verify_ncu.sh ${TAG}_new mnt4_753
SSO_B200_LIB=$PWD/snark-setup-operator_b200/variants/libsso_b200_m4old.so bash tools/gpu_verify_ncu.sh ${TAG}_old mnt4_753
