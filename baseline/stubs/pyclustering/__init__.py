"""Import stub (TEST / BASELINE INFRASTRUCTURE) for pyclustering 0.10.1.2, which the reference's learner imports
(src/learners/homophily_learner.py:6-7) and which is not installed in this image.  PARITY UNPINNED: see cluster/xmeans.py."""
