#!/bin/bash
# tests + bench + launch lists of the base build, then launch lists of the variants in $VARIANTS, then a full capture of $KERNEL
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export TAG=${TAG:-r2o}
bash tools/gpu/r2m.sh
bash tools/gpu/sweep.sh
unset YB_LIB_PATH
[ -n "$KERNEL" ] && bash tools/gpu/prof_kernel.sh
