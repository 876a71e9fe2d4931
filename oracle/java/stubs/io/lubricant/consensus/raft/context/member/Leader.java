package io.lubricant.consensus.raft.context.member;
/** Compile-time stand-in: Membership compares role CLASS TOKENS only (Leader.class). */
public class Leader implements io.lubricant.consensus.raft.RaftParticipant {}
