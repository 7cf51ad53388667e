"""Host-side binary tree used to compose models and parameters (reference: model/Tree.scala).

Only what the filter path needs: the leaf order of `flatten` (model/Tree.scala:49-53) defines the
structure-of-arrays layout of the particle cloud on the device.
"""


class Tree:
    def flatten(self):
        raise NotImplementedError

    def __or__(self, other):  # `a |+| b` of the reference (Semigroup combine) is written `a | b`
        return Branch(self, other)

    def map(self, f):
        raise NotImplementedError

    def getNode(self, n):
        """Leaf value at position n from the left (model/Tree.scala:26-29)."""
        return self.flatten()[n]


class Leaf(Tree):
    def __init__(self, value):
        self.value = value

    def flatten(self):
        return [self.value]

    def map(self, f):
        return Leaf(f(self.value))

    def __repr__(self):
        return f"Leaf({self.value!r})"


class Branch(Tree):
    def __init__(self, left, right):
        self.left, self.right = left, right

    def flatten(self):
        return self.left.flatten() + self.right.flatten()

    def map(self, f):
        return Branch(self.left.map(f), self.right.map(f))

    def __repr__(self):
        return f"Branch({self.left!r}, {self.right!r})"


def zipWith(a, b, f):
    """model/Tree.scala:58-62; trees of different shape raise like the reference."""
    if isinstance(a, Leaf) and isinstance(b, Leaf):
        return Leaf(f(a.value, b.value))
    if isinstance(a, Branch) and isinstance(b, Branch):
        return Branch(zipWith(a.left, b.left, f), zipWith(a.right, b.right, f))
    raise Exception("Can't zip different shaped trees")
