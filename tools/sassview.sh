#!/bin/bash
# strip line markers / encodings from an nvdisasm dump for reading
grep -v "^\s*//##" "$1" | grep -E "^\s+/\*[0-9a-f]{4}\*/|^\.L" | sed -e 's/\/\*[0-9a-f]\{4\}\*\///' -e 's/^\s\+/  /' | awk '{print NR": "$0}'
