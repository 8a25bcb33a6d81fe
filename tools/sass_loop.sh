#!/bin/bash
# SASS of one kernel from build/trace_probe.o (after tools/ptxas_probe.sh), stripped to address + instruction
cuobjdump -sass build/trace_probe.o | awk -v k="$1" '/Function :/{f=index($0,k)>0} f' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed 's#/\* 0x[0-9a-f]* \*/##' | cut -c1-120
