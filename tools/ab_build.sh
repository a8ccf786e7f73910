#!/bin/bash
# Build the library of another git revision next to the current one (videomv_b200/lib/libvideomv_b200_<name>.so) for
# same-box A/B timing:  tools/ab_build.sh HEAD prev ;  VMV_LIB=videomv_b200/lib/libvideomv_b200_prev.so python tools/...
set -e
rev=${1:-HEAD}; name=${2:-prev}
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
git -C "$root" worktree add --detach "$tmp/wt" "$rev" > /dev/null
( cd "$tmp/wt" && python -c "from videomv_b200 import _lib; print(_lib.build(force=True))" )
cp "$tmp/wt/videomv_b200/lib/libvideomv_b200.so" "$root/videomv_b200/lib/libvideomv_b200_${name}.so"
git -C "$root" worktree remove --force "$tmp/wt"
echo "built $root/videomv_b200/lib/libvideomv_b200_${name}.so from $rev"
