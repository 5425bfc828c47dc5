#!/bin/bash
# registers / shared / SASS instruction count of one kernel in each given library
# usage: tools/sass_info.sh <mangled-name-substring> lib1.so lib2.so ...
k=$1; shift
for f in "$@"; do
  n=$(cuobjdump -sass $f | awk -v k="$k" '/Function :/{f=index($0,k)>0;next} f' | grep -cE "^\s+/\*[0-9a-f]{4}\*/")
  r=$(cuobjdump -res-usage $f 2>/dev/null | grep -A1 "$k" | grep -o "REG:[0-9]*\|SHARED:[0-9]*\|STACK:[0-9]*" | tr '\n' ' ')
  echo "$(basename $f) sass=$n $r"
done
