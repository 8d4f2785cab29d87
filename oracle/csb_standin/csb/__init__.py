"""Minimal stand-in for the third-party CSB toolbox (csb-toolbox/CSB, unpinned in the
reference's setup.py:25).  TEST INFRASTRUCTURE ONLY: it exists so that the unmodified
reference package under /root/reference can be imported by oracle/ref_import.py in order
to generate and check golden vectors.  It is not imported by the product package.

Only the names the reference touches are provided."""
