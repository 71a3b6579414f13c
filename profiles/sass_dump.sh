#!/bin/bash
# SASS listings of the tcgen05 kernels of the in-tree library (runs anywhere: cuobjdump needs no GPU) + the counts of the
# Blackwell-native instructions (B200_PROFILING.md: UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit)
#   bash profiles/sass_dump.sh [tag]
TAG=${1:-r02}
LIB=nerf_atlas_b200/libnerf_b200.so
dump() {  # $1 = mangled-name regex, $2 = output
  cuobjdump -sass $LIB | awk -v pat="$1" '/Function :/ { on = ($0 ~ pat) } on' | sed -E 's@/\* 0x[0-9a-f]+ \*/@@; s@^[[:space:]]+/\*[0-9a-f]{4,5}\*/@@; s/[[:space:]]+$//' | grep -v '^$' > $2
}
dump 'k_render_tc3ILi3ELi4ELi4ELi0ELb0ELb0ELb0ELb1E' profiles/${TAG}_k_render_tc3.sass
dump 'k_render_tc3ILi3ELi4ELi4ELi0ELb0ELb1ELb0ELb1E' profiles/${TAG}_k_render_tc3_train.sass
dump 'k_bwd_chain' profiles/${TAG}_k_bwd_chain.sass
dump 'k_bwd_dw' profiles/${TAG}_k_bwd_dw.sass
{
  echo "# SASS instruction counts (static), libnerf_b200.so, $(date -u +%F)"
  for f in profiles/${TAG}_k_render_tc3.sass profiles/${TAG}_k_render_tc3_train.sass profiles/${TAG}_k_bwd_chain.sass profiles/${TAG}_k_bwd_dw.sass; do
    echo "$f: $(wc -l < $f) lines"
    for op in UTCHMMA.2CTA 'UTCHMMA ' LDTM UTCBAR UBLKCP UTMALDG MUFU.SIN MUFU.COS '[^C]HMMA' 'SYNCS.PHASECHK' ATOMG RED STG; do
      printf "  %-16s %s\n" "$op" "$(grep -c -- "$op" $f)"
    done
  done
} > profiles/${TAG}_sass_counts.txt
cat profiles/${TAG}_sass_counts.txt
