for d in 0 64 1 2 3; done
