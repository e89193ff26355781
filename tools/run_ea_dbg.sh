out=$(WB_EA_DBG=148 timeout 200 python tools/ea_dbg.py 64 2>&1 | head -1); echo "$out"
id=$(echo "$out" | sed 's/.*SM://' | awk '{print $1}'); echo "second CTA: $id"
WB_EA_DBG=$id timeout 200 python tools/ea_dbg.py 64 2>&1 | tail -28
