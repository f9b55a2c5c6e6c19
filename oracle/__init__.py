"""CPU oracle (test infrastructure only). See oracle/ref_akaze.c. Never imported by cv_b200/."""
