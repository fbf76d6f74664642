"""Host-side mirror of the reference's ``net/`` package: same class names, constructor / forward
signatures and ``state_dict`` keys (net/model.py:17,31 resolve them by name), with the arithmetic in
libfreqair.so.  PyTorch only owns memory, streams and the autograd graph of coarse block-level nodes."""
