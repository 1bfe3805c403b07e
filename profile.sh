#!/bin/bash
# profile.sh — Nsight Compute capture of the harness, like the reference's profile.sh:1-21
# (ncu --set full with source import on ./build/sparse_sgemv -> ./profile/gemv.ncu-rep).
# No sudo and no hard-wired GPU index: pick the device with CUDA_VISIBLE_DEVICES.
set -euo pipefail
EXE=${EXE:-./build/sparse_sgemv}
OUT_DIR=${OUT_DIR:-./profile}
REPORT=${OUT_DIR}/gemv.ncu-rep
NCU=${NCU:-ncu}
[ -x "${EXE}" ] || { echo "build first: make harness" >&2; exit 1; }
mkdir -p "${OUT_DIR}"
SPMV_SEED=${SPMV_SEED:-1234} "${EXE}" > "${OUT_DIR}/plain_run.log"      # must pass before profiling
SPMV_SEED=${SPMV_SEED:-1234} "${NCU}" --import-source yes --clock-control none --set full -f -o "${REPORT}" "${EXE}"
echo "report: ${REPORT}"
