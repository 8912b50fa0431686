"""oracle — CPU restatement of the reference's shallow-water step.  TEST INFRASTRUCTURE:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this package.  The product (terrainwatersim_b200, libtws.so) never does."""
