#!/bin/bash
# D2H bytes of a 200-query run with a significance cut (-z): the kernels stream the hits, only those travel to the host.
cd "$(dirname "$0")/../.."
python - <<'PY'
import sys
import numpy as np
sys.path.insert(0, ".")
import cuda_satabsearch_b200 as S
base = S.Database.read_packed("tests/golden/small586.satsdb")
db15 = base.bootstrap(14297, 20240501, True)
db15.write_packed("/tmp/db15.satsdb")
rng = np.random.default_rng(200)
open("/tmp/ids200", "w").write("\n".join(db15.name(int(i)) for i in rng.choice(len(db15), 200, replace=False)) + "\n")
PY
for z in 0.5 1.0 2.0; do
  echo "== -z $z"
  cuda_satabsearch_b200/bin/cudaSaTabsearch -q /tmp/db15.satsdb -r 128 -z $z < /tmp/ids200 2>&1 >/tmp/out_z | grep -E "streamed|GPU execution"
  grep -vc "^#" /tmp/out_z
done
