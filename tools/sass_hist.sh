#!/bin/bash
# sass_hist.sh <object> <mangled-name-regex> : opcode histogram of one kernel (cuobjdump -sass)
cuobjdump -sass "$1" 2>/dev/null | awk -v pat="$2" '/Function :/{on = ($3 ~ pat)} on && /^[ \t]+\/\*[0-9a-f][0-9a-f][0-9a-f][0-9a-f]\*\//{print}' |
  sed -E 's/^[ \t]+\/\*[0-9a-f]+\*\/[ \t]+//' | sed -E 's/^@!?U?P[0-9T]+ //' | awk '{print $1}' | sed -E 's/;$//' | awk -F. '{print $1}' | sort | uniq -c | sort -rn
