class _OutOfScope(object):
    def __init__(self, *a, **k):
        raise NotImplementedError("the quaternion model family is out of scope of the B200 hot path (SURVEY.md §2); "
                                  "use --type undirectional-phm-sc-add")


class QuaternionSkipConnectAdd(_OutOfScope):
    pass


class QuaternionSkipConnectConcat(_OutOfScope):
    pass
