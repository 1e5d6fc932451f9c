#!/bin/bash
# Registers / stack (spill) / static shared memory per kernel of the built library, read from the objects with cuobjdump
# (no GPU needed): tools/resource_usage.sh > profiles/rNN_resource_usage.txt
cd "$(dirname "$0")/../ligero-prover_b200/build" || exit 1
echo "# cuobjdump --dump-resource-usage over ligero-prover_b200/build/*.o (sm_100a); STACK > 0 = spills or local arrays"
for f in *.o; do
  cuobjdump --dump-resource-usage "$f" 2>/dev/null | grep -A1 " Function " | paste - - | sed 's/^ *Function //; s/TEXTURE.*//' | while read -r name rest; do
    printf "%-24s %-110s %s\n" "$f" "$(echo "${name%:}" | c++filt | cut -c1-108)" "$rest"
  done
done
