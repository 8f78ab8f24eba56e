#!/bin/bash
# Start-up trace of `cudaSaTabsearch -g N` (SATS_TRACE=1 prints the phases of every searcher's creation): run on a GPU box.
set -e
cd "$(dirname "$0")/../.."
python - <<'PY'
import sys
sys.path.insert(0, ".")
import cuda_satabsearch_b200 as S
base = S.Database.read_packed("tests/golden/small586.satsdb")
qs = S.Database.read_packed("tests/golden/queries.satsdb")
base.bootstrap(100000, 20240502, True).write_packed("/tmp/db100.satsdb")
qs.select([qs.find("D2PHLB1")]).write_ascii("/tmp/q.ascii")
open("/tmp/in100", "w").write("/tmp/db100.satsdb\nT T F\n" + open("/tmp/q.ascii").read())
PY
for g in ${1:-1 2}; do
  for rep in 1 2; do
    echo "== -g $g (run $rep)"
    SATS_TRACE=1 cuda_satabsearch_b200/bin/cudaSaTabsearch -r 128 -g $g < /tmp/in100 2>&1 >/dev/null | grep -v "^Tableau\|WARNING"
  done
done
