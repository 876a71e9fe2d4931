// Upstream golden vectors for include/rafting_ingest.h's reply-body codec (test infrastructure; NOT run in this project's
// images: no JDK, and kryo 4.0.2 — pom.xml:22-26 — is not in the reference tree).  On a box with a JDK, the reference built
// by maven (target/classes) and its dependency jars (kryo-4.0.2, objenesis, minlog, reflectasm, asm, netty-buffer):
//
//     make -C oracle/java kryo REF=/path/to/rafting CP=/path/to/deps/'*'
//
// prints, for a fixed list of (term, success), the bytes the reference's OWN Serialization.writeObject(RaftResponse) emits
// (support/serial/Serialization.java:38-61) as JSON: tests/golden/upstream_kryo.json.  tests/test_ingest_cpu.py replays them
// through rafting_reply_body_encode / _decode and skips while the file is absent.
import io.lubricant.consensus.raft.RaftResponse;
import io.lubricant.consensus.raft.support.serial.Serialization;

public class KryoGen {
    public static void main(String[] a) throws Exception {
        long[] terms = {0, 1, -1, 5, 63, 64, -64, -65, 127, 128, 1L << 31, 1L << 55, (1L << 56) - 1, 1L << 62, Long.MAX_VALUE, Long.MIN_VALUE};
        StringBuilder sb = new StringBuilder("{\"source\": \"Serialization.writeObject(RaftResponse.reply(term, success)), kryo 4.0.2\", \"vectors\": [");
        boolean first = true;
        for (long t : terms) for (int ok = 0; ok < 2; ok++) {
            byte[] b = Serialization.writeObject(RaftResponse.reply(t, ok == 1));
            StringBuilder hex = new StringBuilder();
            for (byte x : b) hex.append(String.format("%02x", x & 0xff));
            if (!first) sb.append(", ");
            first = false;
            sb.append("{\"term\": ").append(t).append(", \"success\": ").append(ok == 1).append(", \"hex\": \"").append(hex).append("\"}");
        }
        System.out.println(sb.append("]}"));
    }
}
