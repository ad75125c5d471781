class RemoveIsolatedNodes(object):
    def __call__(self, data):
        return data
