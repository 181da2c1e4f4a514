"""Input recipes of the reference's algebraic examples (ALG-NB cells 3, 9, 15, 21;
predictions-example1.ipynb cell 1), regenerated with the oracle's JAX-PRNG
restatement.  Test helper only."""
import numpy as np
from oracle import jaxprng


def algebraic_examples(n=1000):
    key = jaxprng.PRNGKey(0)
    out = []
    ks = jaxprng.split(key); key, sub = ks[0], ks[1]
    W = jaxprng.normal(sub, (n, 4))
    out.append((np.concatenate([W[:, :2], W], axis=1),
                ['$x_1$', '$x_2$', '$w_1$', '$w_2$', '$w_3$', '$w_4$']))
    ks = jaxprng.split(key); key, sub = ks[0], ks[1]
    W = jaxprng.normal(sub, (n, 4))
    X1 = W[:, 0]; X2 = X1 ** 2 + 1 + 0.1 * W[:, 1]; X3 = W[:, 2]
    out.append((np.concatenate([np.stack([X1, X2, X3], axis=1), W], axis=1),
                ['$x_1$', '$x_2$', '$x_3$', '$w_1$', '$w_2$', '$w_3$', '$w_4$']))
    ks = jaxprng.split(key); key, sub = ks[0], ks[1]
    W = jaxprng.normal(sub, (n, 4))
    X1 = W[:, 0] * W[:, 1]; X2 = W[:, 1] * np.sin(W[:, 3])
    out.append((np.concatenate([np.stack([X1, X2], axis=1), W], axis=1),
                ['$x_1$', '$x_2$', '$w_1$', '$w_2$', '$w_3$', '$w_4$']))
    ks = jaxprng.split(key); key, sub = ks[0], ks[1]
    W = jaxprng.normal(sub, (n, 4))
    X1 = W[:, 0]; X2 = X1 ** 3 + 1 + 0.1 * W[:, 1]; X3 = (X1 + 2) ** 3 + 0.1 * W[:, 2]
    out.append((np.concatenate([np.stack([X1, X2, X3], axis=1), W], axis=1),
                ['$x_1$', '$x_2$', '$x_3$', '$w_1$', '$w_2$', '$w_3$', '$w_4$']))
    return out


def predictions_example1():
    key = jaxprng.PRNGKey(0)
    ks = jaxprng.split(key)
    W = jaxprng.normal(ks[1], (420, 4))
    return (np.concatenate([W[:, :2], W], axis=1),
            ['$x_1$', '$x_2$', '$w_1$', '$w_2$', '$w_3$', '$w_4$'])
