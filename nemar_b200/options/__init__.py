"""Command-line surface of the reference (options/{base,train,test}_options.py), table-driven."""
