#!/bin/bash
for i in 1 2 3 4 5 6; do ./tests/cpp/_bin/test_bfv_addition | tail -3; done
