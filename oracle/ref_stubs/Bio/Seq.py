class Seq(str):
    def reverse_complement(self):
        comp = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N", "a": "t", "c": "g", "g": "c", "t": "a", "n": "n"}
        return Seq("".join(comp.get(c, c) for c in reversed(self)))
