"""CPU oracle package -- TEST INFRASTRUCTURE ONLY (see oracle/metaseg_oracle.py header).
Nothing under ecseg_b200/ imports from here."""
