"""TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's token-reduction path (oracle.ops,
oracle.model) plus the timm shim used to import the real reference in the build container.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this."""
