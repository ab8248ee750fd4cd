#!/bin/bash
# SASS evidence (runs without a GPU): per kernel of the product library, how many tcgen05 / TMEM / legacy-MMA / reduction
# instructions it contains.  Usage: bash profiles/sass_excerpt.sh > profiles/r02_sass_excerpt.txt
lib=g-nerf_b200/lib/libtriplane_b200.so
echo "# cuobjdump -sass $lib  (sm_100a), instruction counts per kernel"
echo "# UTCHMMA/UTCQMMA... = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM/STTM = tcgen05.ld/st, HMMA = mma.sync, RED = REDG (red.global, scalar or .128), LDGSTS = cp.async, UBLKCP = cp.async.bulk (1-D TMA; the opt-in input prefetch)"
cuobjdump -sass $lib | awk '
  /Function :/ { name=$3; order[++n]=name }
  { for (k in pat) if ($0 ~ pat[k]) cnt[name,k]++ }
  BEGIN { pat["UTCHMMA"]="UTC[A-Z]*MMA"; pat["UTCBAR"]="UTCBAR"; pat["LDTM"]="LDTM"; pat["STTM"]="STTM"; pat["HMMA"]=" HMMA"; pat["RED"]="REDG"; pat["LDGSTS"]="LDGSTS"; pat["SYNCS"]="SYNCS"; pat["UBLKCP"]="UBLKCP" }
  END { printf "%-9s %-7s %-6s %-6s %-6s %-6s %-7s %-6s %-7s %s\n","UTC*MMA","UTCBAR","LDTM","STTM","HMMA","RED","LDGSTS","SYNCS","UBLKCP","kernel";
        for (i=1;i<=n;i++) { k=order[i]; printf "%-9d %-7d %-6d %-6d %-6d %-6d %-7d %-6d %-7d %s\n", cnt[k,"UTCHMMA"],cnt[k,"UTCBAR"],cnt[k,"LDTM"],cnt[k,"STTM"],cnt[k,"HMMA"],cnt[k,"RED"],cnt[k,"LDGSTS"],cnt[k,"SYNCS"],cnt[k,"UBLKCP"], k } }' | (read h1; echo "$h1"; cat) | awk 'NR<=1 || ($1+$2+$3+$4+$5+$6+$7)>0' | c++filt | cut -c1-230
