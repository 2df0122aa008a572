#!/bin/bash
python tests/dev_make_corpus.py c2 2000 /dev/shm/c2.bin
ATZ_DEBUG_SCAN=1 antiz_b200/uncomp -i /dev/shm/c2.bin --notest --stats 2>&1 | tail -30
