class _Comm:
    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1


COMM_WORLD = _Comm()
COMM_SELF = _Comm()
SUM = "sum"
