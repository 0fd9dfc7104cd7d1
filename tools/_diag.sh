for i in 1 2; do timeout 300 python tools/_n3.py 2>&1 | grep "F="; done
