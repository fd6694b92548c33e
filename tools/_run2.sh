B200RT_REFERENCE_ROOT=$PWD/.ab/reference python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -15
