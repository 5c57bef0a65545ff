# usage: tools/sass_equal.sh <git-rev>    -- is the device code of the working tree the same as at <git-rev>?
# Builds <git-rev> in a temporary worktree and compares the SASS instruction streams of the two liborbb200.so (addresses
# and encodings stripped of nothing but the line-info-dependent comments).  Used to accept comment-only / host-only edits
# of csrc/ when no GPU time is left to re-run the parity tests: identical SASS = the kernels that were verified.
set -e
REV=${1:?git revision}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
WT=$(mktemp -d /tmp/sasswt.XXXXXX)
dump() { cuobjdump -sass "$1" | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed 's#/\* 0x[0-9a-f]* \*/##'; }
git -C "$ROOT" worktree add -q --detach "$WT" "$REV"
( cd "$WT" && python vi-orb-slam-icra2018_b200/build.py > /dev/null )
python "$ROOT/vi-orb-slam-icra2018_b200/build.py" > /dev/null
dump "$WT/vi-orb-slam-icra2018_b200/liborbb200.so" > "$WT/old.sass"
dump "$ROOT/vi-orb-slam-icra2018_b200/liborbb200.so" > "$WT/new.sass"
if cmp -s "$WT/old.sass" "$WT/new.sass"; then echo "SASS identical to $REV ($(wc -l < "$WT/new.sass") instructions)"; RC=0
else echo "SASS DIFFERS from $REV"; diff "$WT/old.sass" "$WT/new.sass" | head -20; RC=1; fi
git -C "$ROOT" worktree remove --force "$WT"
exit $RC
