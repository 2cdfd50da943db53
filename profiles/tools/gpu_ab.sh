#!/bin/bash
for tag in base w2b9 w2b10 w3b6; do
  echo "=== $tag"
  if [ $tag = base ]; then python profiles/tools/class_times.py 300 2>&1 | grep -E "main|total"; else CTK_LIB_PATH=/root/repo/profiles/tools/_build/libctk_$tag.so python profiles/tools/class_times.py 300 2>&1 | grep -E "main|total"; fi
done
