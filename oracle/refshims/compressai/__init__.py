"""Stand-in for the four pieces of compressai 1.1.8 the reference touches
(graphs/layers/entropy_layer_nets.py:5,9; graphs/models/LLICTI_nets.py:5,8).
Test infrastructure only."""
