python -m pytest tests -x -q -m gpu -k "sourcewise or source_wise" 2>&1 | tail -30
