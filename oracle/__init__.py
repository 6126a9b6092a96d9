"""CPU oracle (test infrastructure). See physim_oracle.cpp. Not importable from physim_b200/."""
