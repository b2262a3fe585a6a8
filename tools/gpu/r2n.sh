#!/bin/bash
# tests + bench + launch lists, then full ncu captures of the ordering and sorting kernels (C3)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export TAG=${TAG:-r2n}
bash tools/gpu/r2m.sh
KERNEL=order_kernel SKIP=2 bash tools/gpu/prof_kernel.sh
KERNEL=sort_kernel SKIP=2 bash tools/gpu/prof_kernel.sh
