#!/bin/bash
# all GPU tests, bench C3/C5 + launch lists, full capture of bigscan_kernel on C5
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export TAG=${TAG:-r2w}
bash tools/gpu/r2m.sh
WL=c5 KERNEL=bigscan_kernel SKIP=2 bash tools/gpu/prof_kernel.sh
