package io.lubricant.consensus.raft;
/** Compile-time stand-in for the reference's RaftParticipant (src/main/java/io/lubricant/consensus/raft/RaftParticipant.java):
 *  Membership only uses it as the bound of a Class token.  NOT a copy: the real interface declares the RPC handlers. */
public interface RaftParticipant {}
