class SeqRecord(object):
    def __init__(self, seq, id="", name="", description=""):
        self.seq, self.id, self.name, self.description = seq, id, name, description
