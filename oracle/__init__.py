"""oracle/ — TEST INFRASTRUCTURE ONLY.  CPU restatements of the reference hot path used as the parity checker by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  Nothing under
insar-unet-ca_b200/ or unetca_b200/ imports this package."""
