# NumPy>=2 removed np.infty; the reference evaluates it at import time
# (bgflow/distribution/normal.py:126).  Only used when importing the reference
# in the build container (golden generation / oracle validation).
import numpy
if not hasattr(numpy, "infty"):
    numpy.infty = numpy.inf
