#!/bin/bash
# development tool: run every _bin/mb_* at 2^24 elements with the model's default attributes
cd "$(dirname "$0")"
for f in _bin/mb_*; do
  m=$(echo $f | sed 's/.*mb_M_\([A-Za-z]*\)_.*/\1/')
  timeout 60 $f 24 32 _bin/$m.attrs
done
