#!/bin/bash
# Static SASS evidence per kernel of the built library: profiles/sass_evidence.sh > profiles/r02_sass_evidence.txt
LIB=${1:-nerfstudio_thermal_b200/lib/libtn_b200.so}
echo "# cuobjdump -sass $LIB: per kernel, static counts of UTCHMMA (tcgen05.mma) LDTM (tcgen05.ld) UTCBAR (tcgen05.commit) HMMA (mma.sync) REDG (red.global) ELECT BRA.U.ANY (election loops)"
cuobjdump -sass "$LIB" | awk '
/Function :/ { if (name != "") emit(); name=$3; delete c }
/UTCHMMA/ {c["UTCHMMA"]++} /LDTM/ {c["LDTM"]++} /UTCBAR/ {c["UTCBAR"]++} / HMMA/ {c["HMMA"]++} /REDG/ {c["REDG"]++} / ELECT/ {c["ELECT"]++} /BRA.U.ANY/ {c["BRAUANY"]++}
function emit() { if (c["UTCHMMA"]+c["LDTM"]+c["HMMA"]+c["REDG"] > 0) printf "%-110s UTCHMMA %4d LDTM %4d UTCBAR %3d HMMA %3d REDG %3d ELECT %3d BRA.U.ANY %3d\n", substr(name,1,110), c["UTCHMMA"], c["LDTM"], c["UTCBAR"], c["HMMA"], c["REDG"], c["ELECT"], c["BRAUANY"] }
END { emit() }' | c++filt | sort
