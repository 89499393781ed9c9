// TEST INFRASTRUCTURE ONLY -- CPU oracle, SNAP part (filled in below).
