#!/bin/bash
# sass_of.sh <lib.so> <kernel-name-substring> : compact SASS listing (address, instruction) of the first matching kernel
cuobjdump -sass "$1" | awk -v pat="$2" '/Function :/{on=index($0,pat)>0 && !done; if(on) done=1} on' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*([0-9a-f]{4})\*\/\s+/\1 /; s/\s*\/\*.*$//'
