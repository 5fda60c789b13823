#!/bin/sh
# apply.sh <ggdmc checkout> <ggdmc_b200 repository>: see README.md in this directory.
set -eu
PKG=${1:?usage: apply.sh <ggdmc checkout> <ggdmc_b200 repository>}
B200=${2:?usage: apply.sh <ggdmc checkout> <ggdmc_b200 repository>}
HERE=$(cd "$(dirname "$0")" && pwd)
for f in "$PKG/src/RcppExports.cpp" "$PKG/R/RcppExports.R" "$PKG/R/sampling.R" "$PKG/DESCRIPTION"; do
    [ -f "$f" ] || { echo "apply.sh: $f not found -- is $PKG a ggdmc checkout?" >&2; exit 1; }
done
grep -q '_ggdmc_run_batch' "$PKG/src/RcppExports.cpp" && { echo "apply.sh: already applied" >&2; exit 1; }

# 1. the glue replaces the CPU sampler
rm -f "$PKG/src/de.cpp" "$PKG/src/de.h" "$PKG/src/de2R.cpp" "$PKG/src/type_casting.h" "$PKG"/src/*.o
cp "$HERE/../ggdmc_b200_glue.cpp" "$PKG/src/ggdmc_b200_glue.cpp"
# 2. Makevars with the repository's path filled in
sed "s|\$(GGDMC_B200)|$B200|g" "$HERE/Makevars" > "$PKG/src/Makevars"
# 3. wrappers of the added routines: before the CallEntries table, and their rows in it; Armadillo is gone
awk -v block="$HERE/RcppExports_batch.cpp" -v rows="$HERE/CallEntries_batch.inc" '
    /^#include <RcppArmadillo.h>/ { next }
    /^static const R_CallMethodDef CallEntries\[\]/ { while ((getline l < block) > 0) print l }
    /^    \{NULL, NULL, 0\}/ { while ((getline l < rows) > 0) print l }
    { print }' "$PKG/src/RcppExports.cpp" > "$PKG/src/RcppExports.cpp.new"
mv "$PKG/src/RcppExports.cpp.new" "$PKG/src/RcppExports.cpp"
# 4. their R-side stubs
cat "$HERE/RcppExports_batch.R" >> "$PKG/R/RcppExports.R"
grep -q '^export(run_batch)' "$PKG/NAMESPACE" 2>/dev/null || printf 'export(run_batch)\nexport(run_subject_batch)\nexport(sumloglike_init_batch)\nexport(sumlogprior_batch)\n' >> "$PKG/NAMESPACE"
# 5. parallel_lapply: from its definition to the closing brace in column 0
awk -v repl="$HERE/parallel_lapply.R" '
    /^parallel_lapply <- function\(/ { while ((getline l < repl) > 0) print l; skip = 1; next }
    skip && /^}/ { skip = 0; next }
    !skip { print }' "$PKG/R/sampling.R" > "$PKG/R/sampling.R.new"
mv "$PKG/R/sampling.R.new" "$PKG/R/sampling.R"
# 6. LinkingTo
awk '
    /^LinkingTo:/ { print; print "    Rcpp (>= 1.0.7)"; print "SystemRequirements: libggdmc_b200.so (CUDA 12.9, sm_100a), see src/Makevars"; skip = 1; next }
    skip && /^[A-Za-z]/ { skip = 0 }
    !skip { print }' "$PKG/DESCRIPTION" > "$PKG/DESCRIPTION.new"
mv "$PKG/DESCRIPTION.new" "$PKG/DESCRIPTION"
echo "apply.sh: $PKG now builds against $B200 (R CMD INSTALL $PKG)"
