"""`python -m weightedld_b200 --file X [--min-acgt A] [--min-variability V] [--unweighted]` — the
command line of the reference's WeightedLD.py (WeightedLD.py:405-418) on the B200 library: same
flags, same stdout (`posa posb D D' R2`)."""
from .pycompat import build_parser, main

if __name__ == "__main__":
    main(build_parser().parse_args())
