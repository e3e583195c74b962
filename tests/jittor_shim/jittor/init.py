"""unused initialisers (only referenced from constructors that the fixtures never call)"""
